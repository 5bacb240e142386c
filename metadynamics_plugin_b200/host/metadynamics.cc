// metadynamics.cc -- host classes of the reference's operator surface over the C ABI (see metadynamics.h).
#include "metadynamics.h"

#include <sys/stat.h>

#include <iomanip>

namespace metadynamics {

static cudaStream_t stream_of(const std::shared_ptr<ExecutionConfiguration>& e) { return e->stream; }

// ================================================================================================ CollectiveVariable
CollectiveVariable::CollectiveVariable(std::shared_ptr<SystemDefinition> sysdef, const std::string& name)
    : ForceCompute(sysdef), m_cv_name(name), m_d_scalars(4) {}

// host scalars travel to the device as kernel arguments (metad_set_double): no staging buffer, no synchronisation -- a CV
// that is a host scalar (AspectRatio, Density) no longer stalls the device-resident step every time step
const double* CollectiveVariable::getCurrentValueDevice(unsigned int timestep) {
    metad_check(metad_set_double(m_d_scalars.data() + 2, (double)getCurrentValue(timestep), stream_of(m_exec_conf)), "metad_set_double");
    return m_d_scalars.data() + 2;
}

void CollectiveVariable::setBiasFactor(Scalar bias) {
    metad_check(metad_set_double(m_d_scalars.data(), (double)bias, stream_of(m_exec_conf)), "metad_set_double");
}

void CollectiveVariable::setBiasFactorDevice(const double* d_bias) {
    cuda_check(cudaMemcpyAsync(m_d_scalars.data(), d_bias, sizeof(double), cudaMemcpyDeviceToDevice, stream_of(m_exec_conf)), "bias d2d");
}

Scalar CollectiveVariable::getBiasFactor() {
    double b = 0;
    cuda_check(cudaStreamSynchronize(stream_of(m_exec_conf)), "sync");
    cuda_check(cudaMemcpy(&b, m_d_scalars.data(), sizeof(double), cudaMemcpyDeviceToHost), "bias download");
    return (Scalar)b;
}

Scalar CollectiveVariable::biasHost() {
    double b = 0;
    cuda_check(cudaStreamSynchronize(stream_of(m_exec_conf)), "sync");
    cuda_check(cudaMemcpy(&b, biasDevice(), sizeof(double), cudaMemcpyDeviceToHost), "bias download");
    return (Scalar)b;
}

void CollectiveVariable::keepBiasForVirial() {
    cuda_check(cudaMemcpyAsync(m_d_scalars.data() + 3, biasDevice(), sizeof(double), cudaMemcpyDeviceToDevice, stream_of(m_exec_conf)), "bias keep");
    m_virial_dirty = true;
}
Scalar CollectiveVariable::keptBias() {
    double b = 0;
    cuda_check(cudaStreamSynchronize(stream_of(m_exec_conf)), "sync");
    cuda_check(cudaMemcpy(&b, m_d_scalars.data() + 3, sizeof(double), cudaMemcpyDeviceToHost), "bias download");
    return (Scalar)b;
}

// CollectiveVariable.cc:22-66: umbrella increment (evaluated on the device), computeBiasForces, bias reset
void CollectiveVariable::computeForces(unsigned int timestep) {
    m_bias_with_umbrella = false;
    if (m_umbrella != no_umbrella) {
        const double* d_cv = getCurrentValueDevice(timestep);
        metad_check(metad_umbrella_apply((int)m_umbrella, m_cv0, m_kappa, m_width_flat, m_scale, d_cv, m_d_scalars.data(),
                                         m_d_scalars.data() + 1, nullptr, stream_of(m_exec_conf)), "metad_umbrella_apply");
        m_bias_with_umbrella = true;
    }
    computeBiasForces(timestep);
    m_bias_with_umbrella = false;
    cuda_check(cudaMemsetAsync(m_d_scalars.data(), 0, sizeof(double), stream_of(m_exec_conf)), "bias reset");     // setBiasFactor(0.0)
}

// CollectiveVariable.cc:68-106
Scalar CollectiveVariable::getUmbrellaPotential(unsigned int timestep) {
    if (m_umbrella == no_umbrella) return Scalar(0.0);
    const Scalar val = getCurrentValue(timestep);
    if ((val < m_cv0 + m_width_flat / Scalar(2.0)) && (val > m_cv0 - m_width_flat / Scalar(2.0))) return Scalar(0.0);
    Scalar delta(0.0);
    if (val > m_cv0) delta = val - m_cv0 - m_width_flat / Scalar(2.0);
    else if (val < m_cv0) delta = val - m_cv0 + m_width_flat / Scalar(2.0);
    switch (m_umbrella) {
        case linear: return m_scale * delta;
        case harmonic: return Scalar(1.0 / 2.0) * delta * delta * m_kappa;
        case wall: return m_scale * std::pow(delta / m_kappa, Scalar(12.0));
        case gaussian: return m_scale * std::exp(-(val - m_cv0) * (val - m_cv0) / m_kappa / m_kappa / Scalar(2.0)) - m_scale;
        default: return Scalar(0.0);
    }
}

Scalar CollectiveVariable::getLogValue(const std::string& quantity, unsigned int timestep) {
    if (quantity == "umbrella_energy_" + m_cv_name) return getUmbrellaPotential(timestep);
    m_exec_conf->msg->error("cv.*: Invalid log quantity " + quantity);
    throw std::runtime_error("Error querying log quantity");
}

// ================================================================================================ Lamellar
LamellarOrderParameterGPU::LamellarOrderParameterGPU(std::shared_ptr<SystemDefinition> sysdef, const std::vector<Scalar>& mode,
                                                     const std::vector<int3_>& lattice_vectors, const std::string& suffix)
    : CollectiveVariable(sysdef, "cv_lamellar") {
    if (mode.size() != m_pdata->getNTypes()) {
        m_exec_conf->msg->error("cv.lamellar: Number of mode parameters has to equal the number of particle types!");
        throw std::runtime_error("Error initializing cv.lamellar");
    }
    m_cv_name += suffix;
    m_log_name = m_cv_name;
    m_n_wave = (int)lattice_vectors.size();
    std::vector<int> lv;
    for (auto& v : lattice_vectors) { lv.push_back(v.x); lv.push_back(v.y); lv.push_back(v.z); }
    std::vector<double> md(mode.begin(), mode.end());
    metad_check(metad_lamellar_create(&m_plan, m_n_wave, lv.data(), (int)md.size(), md.data()), "Error initializing cv.lamellar");
    m_d_modes.resize(2 * (size_t)m_n_wave);
}
LamellarOrderParameterGPU::~LamellarOrderParameterGPU() { metad_lamellar_destroy(m_plan); }

// LamellarOrderParameterGPU.cc:34-96; the CV stays in device memory (m_d_scalars[2])
void LamellarOrderParameterGPU::computeCV(unsigned int timestep) {
    const metad_box box = m_pdata->getGlobalBox().pod();
    double* d_cv = m_d_scalars.data() + 2;
    const bool sharded = (bool)allreduce;
    metad_check(metad_lamellar_modes(m_plan, (const float*)m_pdata->getPositions().data(), m_pdata->getN(), m_pdata->getNGlobal(), &box,
                                     m_d_modes.data(), sharded ? 0 : 1, d_cv, stream_of(m_exec_conf)), "metad_lamellar_modes");
    if (sharded) {
        allreduce(m_d_modes.data(), 2 * m_n_wave);       // reference: MPI_Allreduce, LamellarOrderParameterGPU.cc:70-77
        metad_check(metad_lamellar_finalize(m_plan, m_d_modes.data(), m_pdata->getNGlobal(), d_cv, stream_of(m_exec_conf)),
                    "metad_lamellar_finalize");
    }
    m_cv_last_updated = timestep;
}
const double* LamellarOrderParameterGPU::getCurrentValueDevice(unsigned int timestep) {
    computeCV(timestep);
    return m_d_scalars.data() + 2;
}
Scalar LamellarOrderParameterGPU::getCurrentValue(unsigned int timestep) {
    computeCV(timestep);
    double v = 0;
    cuda_check(cudaStreamSynchronize(stream_of(m_exec_conf)), "sync");
    cuda_check(cudaMemcpy(&v, m_d_scalars.data() + 2, sizeof(double), cudaMemcpyDeviceToHost), "cv download");
    return (Scalar)v;
}
// LamellarOrderParameterGPU.cc:99-132
void LamellarOrderParameterGPU::computeBiasForces(unsigned int timestep) {
    if (m_cv_last_updated < timestep || timestep == 0) computeCV(timestep);
    const metad_box box = m_pdata->getGlobalBox().pod();
    metad_check(metad_lamellar_forces(m_plan, (const float*)m_pdata->getPositions().data(), (float*)m_force.data(), m_pdata->getN(),
                                      m_pdata->getNGlobal(), &box, biasDevice(), stream_of(m_exec_conf)), "metad_lamellar_forces");
}
std::vector<std::string> LamellarOrderParameterGPU::getProvidedLogQuantities() {
    auto l = CollectiveVariable::getProvidedLogQuantities();
    l.push_back(m_log_name);
    return l;
}
Scalar LamellarOrderParameterGPU::getLogValue(const std::string& quantity, unsigned int timestep) {
    if (quantity == m_log_name) return getCurrentValue(timestep);
    return CollectiveVariable::getLogValue(quantity, timestep);
}

// ================================================================================================ Mesh
OrderParameterMeshGPU::OrderParameterMeshGPU(std::shared_ptr<SystemDefinition> sysdef, unsigned int nx, unsigned int ny,
                                             unsigned int nz, std::vector<Scalar> mode, std::vector<int3_> /*zero_modes*/)
    : CollectiveVariable(sysdef, "mesh") {
    if (mode.size() != m_pdata->getNTypes()) {
        m_exec_conf->msg->error("Number of modes unequal number of particle types.");
        throw std::runtime_error("Error setting up cv.mesh");
    }
    std::vector<double> md(mode.begin(), mode.end());
    // zero_modes only feeds a kernel the reference never launches (SURVEY 8a note 1); accepted and ignored
    metad_check(metad_mesh_create(&m_plan, nx, ny, nz, (int)md.size(), md.data()), "Error initializing cv.mesh");
    // the particle arrays of an MD engine keep their addresses: replay the per-step kernel sequence from a CUDA graph
    metad_check(metad_mesh_set(m_plan, 4, 1), "Error initializing cv.mesh");
}
OrderParameterMeshGPU::~OrderParameterMeshGPU() { metad_mesh_destroy(m_plan); }

// OrderParameterMesh.cc:148-189: the table only enters the virial and the (unused) influence function
void OrderParameterMeshGPU::setTable(const std::vector<Scalar>& K, const std::vector<Scalar>& d_K, Scalar kmin, Scalar kmax) {
    if (kmin < 0 || kmax < 0 || kmax <= kmin) {
        std::ostringstream s; s << "cv.mesh kmin, kmax (" << kmin << "," << kmax << ") is invalid";
        m_exec_conf->msg->error(s.str());
        throw std::runtime_error("Error setting up OrderParameterMesh");
    }
    if (K.size() != d_K.size()) {
        m_exec_conf->msg->error("Convolution kernel and derivative have tables of unequal length");
        throw std::runtime_error("Error setting up OrderParameterMesh");
    }
    m_k_min = kmin; m_k_max = kmax; m_table = K; m_table_d = d_K;
    std::vector<double> dk(d_K.begin(), d_K.end());
    metad_check(metad_mesh_set_table(m_plan, dk.data(), (unsigned)dk.size(), (double)kmin, (double)kmax, m_use_table ? 1 : 0), "metad_mesh_set_table");
}
void OrderParameterMeshGPU::setUseTable(bool use_table) {
    m_use_table = use_table;
    metad_check(metad_mesh_set_table(m_plan, nullptr, 0, 0.0, 0.0, use_table ? 1 : 0), "metad_mesh_set_table");
}
// the arg-max of |f_k|^2 and the virial sums are epilogues of the fused z sweep; they are switched on (for good) the first
// time a log quantity or the pressure asks for them, and the current step is re-evaluated with them
void OrderParameterMeshGPU::enableExtras() {
    if (m_extras) return;
    metad_check(metad_mesh_set(m_plan, 13, 1), "metad_mesh_set");
    m_extras = true;
    m_is_first_step = true;                // forces a re-evaluation of the current step
}
void OrderParameterMeshGPU::computeQmax(unsigned int timestep) {
    enableExtras();
    getCurrentValueDevice(timestep);
    if (timestep && m_q_max_last_computed == timestep) return;
    m_q_max_last_computed = timestep;
    double out[12];
    metad_check(metad_mesh_get(m_plan, 10, out), "metad_mesh_get");
    for (int i = 0; i < 3; ++i) m_q_max[i] = (Scalar)out[6 + i];
    m_sq_max = (Scalar)out[9];
}
void OrderParameterMeshGPU::computeVirial() {
    double out[12];
    metad_check(metad_mesh_get(m_plan, 10, out), "metad_mesh_get");
    const Scalar bias = biasHost();
    for (int i = 0; i < 6; ++i) m_external_virial[i] = bias * (Scalar)out[i];
}

const double* OrderParameterMeshGPU::getCurrentValueDevice(unsigned int timestep) {
    if (m_cv_last_updated == timestep && !m_is_first_step) return m_d_scalars.data() + 2;
    const metad_box box = m_pdata->getBox().pod();
    metad_check(metad_mesh_cv(m_plan, (const float*)m_pdata->getPositions().data(), m_pdata->getN(), m_pdata->getNGlobal(), &box,
                              m_d_scalars.data() + 2, stream_of(m_exec_conf)), "metad_mesh_cv");
    m_is_first_step = false;
    m_cv_last_updated = timestep;
    return m_d_scalars.data() + 2;
}
Scalar OrderParameterMeshGPU::getCurrentValue(unsigned int timestep) {
    getCurrentValueDevice(timestep);
    double v = 0;
    cuda_check(cudaStreamSynchronize(stream_of(m_exec_conf)), "sync");
    cuda_check(cudaMemcpy(&v, m_d_scalars.data() + 2, sizeof(double), cudaMemcpyDeviceToHost), "cv download");
    return (Scalar)v;
}
// OrderParameterMesh.cc:1052-1075: forces; the k-space virial only when the pressure is asked for (PDataFlags).  Without a
// kernel table the virial of the reference is identically zero (val_D = 0, :1011-1030): nothing to compute then.
void OrderParameterMeshGPU::computeBiasForces(unsigned int timestep) {
    const bool want_virial = m_pdata->getPressureFlag() && m_use_table && !m_table_d.empty();
    if (want_virial) enableExtras();
    if (m_is_first_step || m_cv_last_updated != timestep) getCurrentValueDevice(timestep);
    const metad_box box = m_pdata->getBox().pod();
    metad_check(metad_mesh_forces(m_plan, (const float*)m_pdata->getPositions().data(), (float*)m_force.data(), m_pdata->getN(),
                                  m_pdata->getNGlobal(), &box, biasDevice(), stream_of(m_exec_conf)), "metad_mesh_forces");
    if (want_virial) computeVirial();
    else for (auto& v : m_external_virial) v = Scalar(0.0);
}
std::vector<std::string> OrderParameterMeshGPU::getProvidedLogQuantities() {
    auto l = CollectiveVariable::getProvidedLogQuantities();
    for (const char* n : {"cv_mesh", "qx_max", "qy_max", "qz_max", "sq_max"}) l.push_back(n);      // OrderParameterMesh.cc:118-122
    return l;
}
Scalar OrderParameterMeshGPU::getLogValue(const std::string& quantity, unsigned int timestep) {
    if (quantity == "cv_mesh") return getCurrentValue(timestep);
    if (quantity == "qx_max") { computeQmax(timestep); return m_q_max[0]; }
    if (quantity == "qy_max") { computeQmax(timestep); return m_q_max[1]; }
    if (quantity == "qz_max") { computeQmax(timestep); return m_q_max[2]; }
    if (quantity == "sq_max") { computeQmax(timestep); return m_sq_max; }
    return CollectiveVariable::getLogValue(quantity, timestep);
}

// ================================================================================================ WTE
WellTemperedEnsemble::WellTemperedEnsemble(std::shared_ptr<SystemDefinition> sysdef, const std::string& name)
    : CollectiveVariable(sysdef, name), m_log_name("cv_potential_energy") {}

// WellTemperedEnsemble.cc:30-68
const double* WellTemperedEnsemble::getCurrentValueDevice(unsigned int) {
    metad_check(metad_wte_reduce((const float*)m_pdata->getNetForce().data(), m_pdata->getN(), m_pdata->getExternalEnergy(),
                                 m_d_scalars.data() + 2, stream_of(m_exec_conf)), "metad_wte_reduce");
    return m_d_scalars.data() + 2;
}
Scalar WellTemperedEnsemble::getCurrentValue(unsigned int timestep) {
    getCurrentValueDevice(timestep);
    double v = 0;
    cuda_check(cudaStreamSynchronize(stream_of(m_exec_conf)), "sync");
    cuda_check(cudaMemcpy(&v, m_d_scalars.data() + 2, sizeof(double), cudaMemcpyDeviceToHost), "pe download");
    return (Scalar)v;
}
// WellTemperedEnsemble.cc:135-188: operates directly on the net force, must run after every other force
void WellTemperedEnsemble::computeBiasForces(unsigned int) {
    metad_check(metad_wte_scale((float*)m_pdata->getNetForce().data(), (float*)m_pdata->getNetTorqueArray().data(),
                                m_pdata->getNetVirial().data(), m_pdata->getNetVirialPitch(), m_pdata->getN(), biasDevice(),
                                stream_of(m_exec_conf)), "metad_wte_scale");
    // external virial of the other forces times the same factor (:180-184).  The factor lives on the device: it is read back
    // only if there is something to scale (no external virial contributions: no host round trip in the step)
    bool any = false;
    for (unsigned int i = 0; i < 6; ++i) any = any || m_pdata->getExternalVirial(i) != Scalar(0.0);
    if (any) {
        const Scalar fac = Scalar(1.0) + biasHost();    // bias incl. the umbrella increment, like the device side
        for (unsigned int i = 0; i < 6; ++i) m_pdata->setExternalVirial(i, fac * m_pdata->getExternalVirial(i));
    }
}
std::vector<std::string> WellTemperedEnsemble::getProvidedLogQuantities() {
    auto l = CollectiveVariable::getProvidedLogQuantities();
    l.push_back(m_log_name);
    return l;
}
Scalar WellTemperedEnsemble::getLogValue(const std::string& quantity, unsigned int timestep) {
    if (quantity == m_log_name) return getCurrentValue(timestep);
    return CollectiveVariable::getLogValue(quantity, timestep);
}

// ================================================================================================ CollectiveWrapper
CollectiveWrapper::CollectiveWrapper(std::shared_ptr<SystemDefinition> sysdef, std::shared_ptr<ForceCompute> fc, const std::string& name)
    : CollectiveVariable(sysdef, name), m_fc(fc), m_d_fac(1) {
    if (!fc) throw std::runtime_error("cv.wrap needs a force to wrap");
}
// CollectiveWrapper.cc:31-73: the wrapped force is computed (once per time step), its per-particle energies are summed,
// its external energy is added, the ranks of a domain decomposition are reduced
const double* CollectiveWrapper::getCurrentValueDevice(unsigned int timestep) {
    m_fc->compute(timestep);
    double* d_cv = m_d_scalars.data() + 2;
    metad_check(metad_wte_reduce((const float*)m_fc->getForceArray().data(), m_pdata->getN(), allreduce ? 0.0 : (double)m_fc->getExternalEnergy(),
                                 d_cv, stream_of(m_exec_conf)), "metad_wte_reduce");
    if (allreduce) {        // local sums first, the external energy of every rank is part of its local sum in the reference (:60-61 before :63-69)
        const double ext = (double)m_fc->getExternalEnergy();
        double local = 0;
        cuda_check(cudaStreamSynchronize(stream_of(m_exec_conf)), "sync");
        cuda_check(cudaMemcpy(&local, d_cv, sizeof(double), cudaMemcpyDeviceToHost), "download");
        local += ext;
        cuda_check(cudaMemcpy(d_cv, &local, sizeof(double), cudaMemcpyHostToDevice), "upload");
        allreduce((size_t)d_cv, 1);
    }
    return d_cv;
}
Scalar CollectiveWrapper::getCurrentValue(unsigned int timestep) {
    getCurrentValueDevice(timestep);
    double v = 0;
    cuda_check(cudaStreamSynchronize(stream_of(m_exec_conf)), "sync");
    cuda_check(cudaMemcpy(&v, m_d_scalars.data() + 2, sizeof(double), cudaMemcpyDeviceToHost), "cv download");
    return (Scalar)v;
}
// CollectiveWrapper.cc:140-188: force, torque (all four components) and virial of the WRAPPED force times the bias factor.
// metad_wte_scale multiplies by (1 + *d_bias) -- hand it bias - 1.
void CollectiveWrapper::computeBiasForces(unsigned int timestep) {
    m_fc->compute(timestep);
    cudaStream_t st = stream_of(m_exec_conf);
    const double fac_minus_one = (double)biasHost() - 1.0;
    cuda_check(cudaMemcpyAsync(m_d_fac.data(), &fac_minus_one, sizeof(double), cudaMemcpyHostToDevice, st), "fac upload");
    cuda_check(cudaStreamSynchronize(st), "fac upload sync");
    metad_check(metad_wte_scale((float*)m_fc->getForceArray().data(), (float*)m_fc->getTorqueArray().data(), m_fc->getVirialArray().data(),
                                (unsigned)m_fc->getVirialPitch(), m_pdata->getN(), m_d_fac.data(), st), "metad_wte_scale");
}

// ================================================================================================ box CVs
AspectRatio::AspectRatio(std::shared_ptr<SystemDefinition> sysdef, unsigned int dir1, unsigned int dir2)
    : CollectiveVariable(sysdef, "cv_aspect_ratio"), m_dir1(dir1), m_dir2(dir2) {
    if (dir1 == dir2 || dir1 >= 3 || dir2 >= 3) {
        m_exec_conf->msg->error("metadynamics.aspect_ratio: Invalid directions given.");
        throw std::runtime_error("Error setting up metadynamics.aspect_ratio");
    }
}
// AspectRatio.cc:24-57 (dir2 == 0 overwrites length1: the reference's bug, kept so results match)
Scalar AspectRatio::getCurrentValue(unsigned int) {
    const Scalar3 L = m_pdata->getGlobalBox().getL();
    const Scalar l[3] = {L.x, L.y, L.z};
    Scalar length1 = l[m_dir1], length2(0.0);
    if (m_dir2 == 0) length1 = L.x;
    else length2 = l[m_dir2];
    return length1 / length2;
}
// AspectRatio.cc:59-130: the "bias force" of a box CV is an external virial, -bias ds/dL_a L_a.  The factor stays on the
// device (it comes out of the grid kernel); the six numbers are evaluated when they are read.
void AspectRatio::computeBiasForces(unsigned int) { keepBiasForVirial(); }
void AspectRatio::updateExternalVirial() {
    if (!m_virial_dirty) return;
    m_virial_dirty = false;
    const BoxDim& box = m_pdata->getGlobalBox();
    const Scalar3 L = box.getL();
    const Scalar bias = keptBias();
    Scalar dx(0.0), dy(0.0), dz(0.0);
    if (m_dir1 == 0 && m_dir2 == 1) { dx = Scalar(1.0) / L.y; dy = -L.x / L.y / L.y; }
    else if (m_dir1 == 0 && m_dir2 == 2) { dx = Scalar(1.0) / L.z; dz = -L.x / L.z / L.z; }
    else if (m_dir1 == 1 && m_dir2 == 0) { dx = -L.y / L.x / L.x; dy = Scalar(1.0) / L.x; }
    else if (m_dir1 == 1 && m_dir2 == 2) { dy = Scalar(1.0) / L.z; dz = -L.y / L.z / L.z; }
    else if (m_dir1 == 2 && m_dir2 == 0) { dx = -L.z / L.x / L.x; dz = Scalar(1.0) / L.x; }
    else if (m_dir1 == 2 && m_dir2 == 1) { dy = -L.z / L.y / L.y; dz = Scalar(1.0) / L.y; }
    m_external_virial[0] = -bias * dx * L.x;
    m_external_virial[1] = -bias * dx * (L.y * (Scalar)box.getTiltFactorXY());
    m_external_virial[2] = -bias * dx * (L.z * (Scalar)box.getTiltFactorXZ());
    m_external_virial[3] = -bias * dy * L.y;
    m_external_virial[4] = -bias * dy * (L.z * (Scalar)box.getTiltFactorYZ());
    m_external_virial[5] = -bias * dz * L.z;
}

Density::Density(std::shared_ptr<SystemDefinition> sysdef, const std::string& suffix)
    : CollectiveVariable(sysdef, "cv_density" + (suffix != "" ? "_" + suffix : "")) {}
// Density.cc:20-27 (group = all particles in the shim)
Scalar Density::getCurrentValue(unsigned int) { return (Scalar)m_pdata->getNGlobal() / (Scalar)m_pdata->getGlobalBox().getVolume(); }
// Density.cc:29-54 (evaluated lazily, like AspectRatio)
void Density::computeBiasForces(unsigned int) { keepBiasForVirial(); }
void Density::updateExternalVirial() {
    if (!m_virial_dirty) return;
    m_virial_dirty = false;
    const BoxDim& box = m_pdata->getGlobalBox();
    const Scalar V = (Scalar)box.getVolume();
    const Scalar3 L = box.getL();
    const Scalar fac = -(Scalar)m_pdata->getNGlobal() / (V * V);
    const Scalar v = -keptBias() * fac * L.x * L.y * L.z;
    m_external_virial[0] = v; m_external_virial[1] = 0; m_external_virial[2] = 0;
    m_external_virial[3] = v; m_external_virial[4] = 0; m_external_virial[5] = v;
}

// ================================================================================================ IndexGrid (IndexGrid.cc)
IndexGrid::IndexGrid() { m_lengths.assign(1, 0); m_factors.assign(1, 1); }
void IndexGrid::setLengths(const std::vector<unsigned int>& lengths) {
    m_lengths = lengths;
    m_factors.resize(lengths.size());
    for (unsigned int i = 0; i < m_lengths.size(); i++) m_factors[i] = (i == 0) ? 1 : (m_lengths[i - 1] * m_factors[i - 1]);
}
unsigned int IndexGrid::getIndex(const std::vector<unsigned int>& coords) {
    unsigned int idx = 0;
    for (unsigned int i = 0; i < m_lengths.size(); i++) idx += coords[i] * m_factors[i];
    return idx;
}
void IndexGrid::getCoordinates(const unsigned int idx, std::vector<unsigned int>& coords) {
    unsigned int rest = idx;
    for (int i = (int)m_lengths.size() - 1; i >= 0; i--) { coords[i] = rest / m_factors[i]; rest -= coords[i] * m_factors[i]; }
}
unsigned int IndexGrid::getNumElements() {
    unsigned int res = 1;
    for (unsigned int l : m_lengths) res *= l;
    return res;
}

// ================================================================================================ IntegratorMetaDynamics
IntegratorMetaDynamics::IntegratorMetaDynamics(std::shared_ptr<SystemDefinition> sysdef, Scalar deltaT, Scalar W, Scalar T_shift,
                                               Scalar T, unsigned int stride, bool add_bias, const std::string& filename,
                                               bool overwrite, const Enum mode)
    : m_sysdef(sysdef), m_pdata(sysdef->getParticleData()), m_exec_conf(sysdef->getExecConf()), m_deltaT(deltaT), m_W(W),
      m_T_shift(T_shift), m_stride(stride), m_filename(filename), m_overwrite(overwrite), m_add_bias(add_bias), m_temp(T), m_mode(mode) {
    m_log_names = {"bias", "det_sigma", "weight"};
}
IntegratorMetaDynamics::~IntegratorMetaDynamics() {
    if (m_grid) metad_grid_destroy(m_grid);
    if (m_h_pinned) cudaFreeHost(m_h_pinned);
}

void IntegratorMetaDynamics::registerCollectiveVariable(std::shared_ptr<CollectiveVariable> cv, Scalar sigma, Scalar cv_min,
                                                        Scalar cv_max, int num_points) {
    CollectiveVariableItem item;
    item.m_cv = cv; item.m_sigma = sigma; item.m_cv_min = cv_min; item.m_cv_max = cv_max; item.m_num_points = (unsigned int)num_points;
    m_variables.push_back(item);
}

// IntegratorMetaDynamics.cc:778-815
void IntegratorMetaDynamics::setGrid(bool use_grid) {
    if (m_is_initialized) {
        m_exec_conf->msg->error("integrate.mode_metadynamics: Cannot change grid mode after initialization.");
        throw std::runtime_error("Error setting up metadynamics parameters.");
    }
    m_use_grid = use_grid;
    if (use_grid)
        for (auto& it : m_variables) {
            if (it.m_cv_min >= it.m_cv_max) {
                m_exec_conf->msg->error("integrate.mode_metadyanmics: Maximum grid value of collective variable has to be greater than minimum value.");
                throw std::runtime_error("Error creating collective variable.");
            }
            if (it.m_num_points < 2) {
                m_exec_conf->msg->error("integrate.mode_metadynamics: Number of grid points for collective variable has to be at least two.");
                throw std::runtime_error("Error creating collective variable.");
            }
        }
}

void IntegratorMetaDynamics::pushFlags() {
    if (m_grid) metad_check(metad_grid_set_flags(m_grid, m_add_bias, m_mode == mode_well_tempered, m_stride), "metad_grid_set_flags");
}
void IntegratorMetaDynamics::setMode(Enum mode) { m_mode = mode; pushFlags(); }
void IntegratorMetaDynamics::setStride(unsigned int stride) { m_stride = stride; pushFlags(); }
void IntegratorMetaDynamics::setAddHills(bool add_bias) { m_add_bias = add_bias; pushFlags(); }
void IntegratorMetaDynamics::setAdaptive(bool adaptive) { m_adaptive = adaptive; }

// computeSigma, IntegratorMetaDynamics.cc:1205-1294: sigma^2_ij = sigma_g^2 sum_n dCV_i/dr_n . dCV_j/dr_n for CVs that can
// compute derivatives (their force arrays hold the derivatives after computeDerivatives), sigma_i^2 on the diagonal
// otherwise; sigma_inv = inverse of the matrix of element-wise square roots (Eigen -> Gauss-Jordan with partial pivoting).
// The N-length sums run on the device (metad_force_dot, fp64); the n_cv x n_cv algebra on the host, on deposit steps only.
void IntegratorMetaDynamics::computeSigma() {
    const size_t d = m_variables.size();
    cudaStream_t st = stream_of(m_exec_conf);
    m_d_sigmasq.resize(d * d);
    cuda_check(cudaMemsetAsync(m_d_sigmasq.data(), 0, sizeof(double) * d * d, st), "sigmasq reset");
    for (size_t i = 0; i < d; ++i)
        for (size_t j = 0; j < d; ++j)
            if (m_variables[i].m_cv->canComputeDerivatives() && m_variables[j].m_cv->canComputeDerivatives())
                metad_check(metad_force_dot((const float*)m_variables[i].m_cv->getForceArray().data(),
                                            (const float*)m_variables[j].m_cv->getForceArray().data(), m_pdata->getN(),
                                            (double)m_sigma_g * (double)m_sigma_g, m_d_sigmasq.data() + i * d + j, st), "metad_force_dot");
    if (domain_allreduce) {
        cuda_check(cudaStreamSynchronize(st), "sync");
        domain_allreduce((size_t)m_d_sigmasq.data(), d * d);
    }
    std::vector<double> sq(d * d);
    cuda_check(cudaStreamSynchronize(st), "sync");
    m_d_sigmasq.download(sq.data(), d * d);
    for (size_t i = 0; i < d; ++i)
        if (!m_variables[i].m_cv->canComputeDerivatives()) sq[i * d + i] = (double)m_variables[i].m_sigma * (double)m_variables[i].m_sigma;
    std::vector<double> m(d * d), inv(d * d, 0.0);
    for (size_t k = 0; k < d * d; ++k) m[k] = std::sqrt(sq[k]);
    for (size_t i = 0; i < d; ++i) inv[i * d + i] = 1.0;
    for (size_t c = 0; c < d; ++c) {
        size_t piv = c;
        for (size_t r = c + 1; r < d; ++r) if (std::fabs(m[r * d + c]) > std::fabs(m[piv * d + c])) piv = r;
        if (piv != c) for (size_t k = 0; k < d; ++k) { std::swap(m[piv * d + k], m[c * d + k]); std::swap(inv[piv * d + k], inv[c * d + k]); }
        const double p = m[c * d + c];
        for (size_t k = 0; k < d; ++k) { m[c * d + k] /= p; inv[c * d + k] /= p; }
        for (size_t r = 0; r < d; ++r) {
            if (r == c) continue;
            const double f = m[r * d + c];
            for (size_t k = 0; k < d; ++k) { m[r * d + k] -= f * m[c * d + k]; inv[r * d + k] -= f * inv[c * d + k]; }
        }
    }
    m_sigma_inv = inv;
    metad_check(metad_grid_set_sigma_inv(m_grid, inv.data()), "metad_grid_set_sigma_inv");
}

// IntegratorMetaDynamics.cc:590-661: the grid arrays live in device memory (metad_grid)
void IntegratorMetaDynamics::setupGrid() {
    const size_t d = m_variables.size();
    std::vector<unsigned int> lengths(d);
    std::vector<double> mn(d), mx(d), sg(d);
    for (size_t i = 0; i < d; ++i) {
        lengths[i] = m_variables[i].m_num_points;
        mn[i] = m_variables[i].m_cv_min; mx[i] = m_variables[i].m_cv_max; sg[i] = m_variables[i].m_sigma;
    }
    m_grid_index.setLengths(lengths);
    if (m_grid) { metad_grid_destroy(m_grid); m_grid = nullptr; }
    metad_check(metad_grid_create(&m_grid, (int)d, mn.data(), mx.data(), lengths.data(), sg.data(), m_W, m_T_shift, m_temp, m_stride,
                                  m_add_bias, m_mode == mode_well_tempered), "Error setting up the bias grid");
}

// IntegratorMetaDynamics.cc:121-217
void IntegratorMetaDynamics::prepRun(unsigned int timestep) {
    if (!m_is_initialized && m_filename != "") {
        openOutputFile();
        if (!m_is_appending) writeFileHeader();
    }
    if (!m_is_initialized) {
        const size_t d = m_variables.size();
        m_d_cv.resize(d ? d : 1);
        m_d_bias.resize(d ? d : 1);
        if (!m_h_pinned) cuda_check(cudaMallocHost(&m_h_pinned, sizeof(double) * 8), "cudaMallocHost");
    }
    if (!m_is_initialized && m_use_grid && m_variables.size()) {
        setupGrid();
        if (m_restart_filename != "") {
            m_exec_conf->msg->notice(2, "integrate.mode_metadynamics: Restarting from grid file \"" + m_restart_filename + "\"");
            readGrid(m_restart_filename);
            m_restart_filename = "";
        }
    }
    m_is_initialized = true;
    updateBiasPotential(timestep);      // initial update of the potential (:213-214)
    m_prepared = true;
}

// IntegratorMetaDynamics.cc:219-312 without integration methods (HOOMD's, out of scope): bias update + net force
void IntegratorMetaDynamics::update(unsigned int timestep) {
    if (!m_prepared) throw std::runtime_error("IntegratorMetaDynamics::update called before prepRun");
    const bool net_force_first = (m_variables.size() == 1 && m_variables[0].m_cv->requiresNetForce());
    if (!net_force_first)
        for (auto& it : m_variables)
            if (it.m_cv->requiresNetForce())
                throw std::runtime_error("Only one collective variable requiring the potential energy may be defined.\n");
    if (net_force_first) computeNetForce(timestep + 1);
    updateBiasPotential(timestep + 1);
    if (!net_force_first) computeNetForce(timestep + 1);
    else m_variables[0].m_cv->compute(timestep);        // bias forces *after* everything else (:297-301)
}

// stand-in for Integrator::computeNetForceGPU: net force = sum of the enabled force computes' force arrays
void IntegratorMetaDynamics::computeNetForce(unsigned int timestep) {
    if (!m_system) return;
    bool first = true;
    for (auto& fc : m_system->computes()) {
        if (!fc->enabled) continue;
        fc->compute(timestep);
        metad_check(metad_accumulate_force((float*)m_pdata->getNetForce().data(), (const float*)fc->getForceArray().data(),
                                           m_pdata->getN(), first ? 1 : 0, stream_of(m_exec_conf)), "metad_accumulate_force");
        first = false;
    }
}

// IntegratorMetaDynamics.cc:314-588, grid mode.  CV values and bias factors stay in device memory.
void IntegratorMetaDynamics::updateBiasPotential(unsigned int timestep) {
    if (m_variables.size() == 0) return;
    if (!m_use_grid) {
        m_exec_conf->msg->error("integrate.mode_metadynamics: only grid mode is supported (the Python API always enables it).");
        throw std::runtime_error("Error in metadynamics integration.");
    }
    const size_t d = m_variables.size();
    cudaStream_t st = stream_of(m_exec_conf);
    for (size_t i = 0; i < d; ++i) {
        const double* d_val = m_variables[i].m_cv->getCurrentValueDevice(timestep);
        cuda_check(cudaMemcpyAsync(m_d_cv.data() + i, d_val, sizeof(double), cudaMemcpyDeviceToDevice, st), "cv gather");
    }
    // adaptive Gaussians (:333-341): derivatives of every CV, then the instantaneous sigma matrix
    if (m_adaptive && (timestep % m_stride == 0)) {
        for (size_t i = 0; i < d; ++i) m_variables[i].m_cv->computeDerivatives(timestep);
        computeSigma();
    }
    if (m_multiple_walkers && walker_allreduce) {
        // multiple walkers (:392-410): the four delta arrays are summed over the walkers between deposit and merge
        metad_check(metad_grid_step_deposit(m_grid, timestep, m_d_cv.data(), st), "metad_grid_step_deposit");
        if (metad_grid_is_deposit_step(m_grid, timestep)) {
            const size_t G = metad_grid_num_elements(m_grid);
            m_d_walk_d.resize(2 * G); m_d_walk_u.resize(2 * G);
            metad_check(metad_grid_deltas_export(m_grid, m_d_walk_d.data(), m_d_walk_u.data(), st), "metad_grid_deltas_export");
            cuda_check(cudaStreamSynchronize(st), "sync");
            walker_allreduce((size_t)m_d_walk_d.data(), 2 * G, (size_t)m_d_walk_u.data(), 2 * G);
            metad_check(metad_grid_deltas_import(m_grid, m_d_walk_d.data(), m_d_walk_u.data(), st), "metad_grid_deltas_import");
        }
        metad_check(metad_grid_step_merge(m_grid, timestep, m_d_cv.data(), m_d_bias.data(), st), "metad_grid_step_merge");
    } else {
        metad_check(metad_grid_step(m_grid, timestep, m_d_cv.data(), m_d_bias.data(), st), "metad_grid_step");
    }

    // hills file (:523-550) -- needs the current bias potential on the host, only when a log file was requested
    if (m_is_initialized && (timestep % m_stride == 0) && m_add_bias && m_file.is_open()) {
        double sc[4];
        metad_check(metad_grid_scalars(m_grid, sc), "metad_grid_scalars");
        std::vector<double> cur(d);
        cuda_check(cudaMemcpy(cur.data(), m_d_cv.data(), sizeof(double) * d, cudaMemcpyDeviceToHost), "cv download");
        const Scalar W = m_W * std::exp(-(Scalar)sc[0] / m_T_shift);
        m_file << std::setprecision(10) << timestep << m_delimiter;
        m_file << std::setprecision(10) << W << m_delimiter;
        for (size_t i = 0; i < d; ++i) {
            m_file << std::setprecision(10) << (Scalar)cur[i] << m_delimiter;
            for (size_t j = 0; j < d; ++j)
                m_file << std::setprecision(10) << (m_sigma_inv.size() == d * d ? (Scalar)m_sigma_inv[i * d + j] : (i == j ? Scalar(1.0) / m_variables[i].m_sigma : Scalar(0.0)));
            if (i != d - 1) m_file << m_delimiter;
        }
        m_file << std::endl;
    }
    // periodic grid dump with the alternating scheme (:555-565)
    if (m_grid_period && (timestep % m_grid_period == 0)) {
        if (m_grid_fname2 != "") {
            writeGrid(m_cur_file ? m_grid_fname2 : m_grid_fname1, timestep);
            m_cur_file = m_cur_file ? 0 : 1;
        } else
            writeGrid(m_grid_fname1, timestep);
    }
    for (size_t i = 0; i < d; ++i) m_variables[i].m_cv->setBiasFactorDevice(m_d_bias.data() + i);
}

Scalar IntegratorMetaDynamics::getLogValue(const std::string& quantity, unsigned int) {
    double sc[4] = {0, 1, 0, 0};
    if (m_grid) metad_check(metad_grid_scalars(m_grid, sc), "metad_grid_scalars");
    if (quantity == m_log_names[0]) return (Scalar)sc[0];
    if (quantity == m_log_names[1]) {
        const size_t d = m_variables.size();
        if (m_sigma_inv.size() == d * d && d) {          // adaptive Gaussians: determinant of the current inverse sigma matrix (sigmaDeterminant :1296-1313)
            std::vector<double> m(m_sigma_inv);
            double det = 1.0;
            for (size_t c = 0; c < d; ++c) {
                size_t piv = c;
                for (size_t r = c + 1; r < d; ++r) if (std::fabs(m[r * d + c]) > std::fabs(m[piv * d + c])) piv = r;
                if (m[piv * d + c] == 0.0) return Scalar(0.0);
                if (piv != c) { for (size_t k = 0; k < d; ++k) std::swap(m[piv * d + k], m[c * d + k]); det = -det; }
                det *= m[c * d + c];
                for (size_t r = c + 1; r < d; ++r) { const double f = m[r * d + c] / m[c * d + c]; for (size_t k = c; k < d; ++k) m[r * d + k] -= f * m[c * d + k]; }
            }
            return (Scalar)det;
        }
        Scalar det = 1;
        for (auto& v : m_variables) det *= Scalar(1.0) / v.m_sigma;
        return det;
    }
    if (quantity == m_log_names[2]) return (Scalar)sc[1];
    std::cerr << std::endl << "***Error! " << quantity << " is not a valid log quantity for IntegratorMetaDynamics" << std::endl << std::endl;
    throw std::runtime_error("Error getting log value");
}

void IntegratorMetaDynamics::resetHistogram() {
    if (m_grid) metad_check(metad_grid_reset_histogram(m_grid, stream_of(m_exec_conf)), "metad_grid_reset_histogram");
}

unsigned int IntegratorMetaDynamics::getNumGaussians() {
    double sc[4] = {0, 0, 0, 0};
    if (m_grid) metad_check(metad_grid_scalars(m_grid, sc), "metad_grid_scalars");
    return (unsigned int)sc[2];
}

std::vector<double> IntegratorMetaDynamics::getGridArray(const std::string& name) {
    static const std::map<std::string, int> ids = {{"grid", 0}, {"reweighted", 1}, {"weight", 2}, {"sigma_grid", 3}, {"hist", 4}, {"hist_gauss", 5}};
    if (!m_grid) throw std::runtime_error("Grid information is only available if the grid is enabled.");
    const int id = ids.at(name);
    const unsigned int G = metad_grid_num_elements(m_grid);
    std::vector<double> out(G);
    if (id < 4) {
        metad_check(metad_grid_download(m_grid, id, out.data()), "metad_grid_download");
    } else {
        std::vector<unsigned int> u(G);
        metad_check(metad_grid_download(m_grid, id, u.data()), "metad_grid_download");
        for (unsigned int i = 0; i < G; ++i) out[i] = u[i];
    }
    return out;
}

// IntegratorMetaDynamics.cc:817-829
void IntegratorMetaDynamics::dumpGrid(const std::string& filename1, const std::string& filename2, unsigned int period) {
    if (period == 0) { writeGrid(filename1, 0); return; }
    m_grid_period = period; m_grid_fname1 = filename1; m_grid_fname2 = filename2;
}

// IntegratorMetaDynamics.cc:831-926 -- same text format (header lines, column order, setprecision(10))
void IntegratorMetaDynamics::writeGrid(const std::string& filename, unsigned int timestep) {
    if (!m_use_grid || !m_grid) {
        m_exec_conf->msg->error("integrate.mode_metadynamics: Grid information can only be dumped if grid is enabled.");
        throw std::runtime_error("Error dumping grid.");
    }
    const unsigned int len = m_grid_index.getNumElements();
    std::vector<double> grid(len), sigma(len), rew(len), weight(len);
    std::vector<unsigned int> hist(len), hist_gauss(len);
    metad_check(metad_grid_download(m_grid, 0, grid.data()), "grid download");
    metad_check(metad_grid_download(m_grid, 1, rew.data()), "grid download");
    metad_check(metad_grid_download(m_grid, 2, weight.data()), "grid download");
    metad_check(metad_grid_download(m_grid, 3, sigma.data()), "grid download");
    metad_check(metad_grid_download(m_grid, 4, hist.data()), "grid download");
    metad_check(metad_grid_download(m_grid, 5, hist_gauss.data()), "grid download");

    std::ofstream file((filename + "_" + std::to_string(timestep)).c_str(), std::ios_base::out);
    file << "#n_cv: " << m_grid_index.getDimension() << std::endl;
    file << "#dim: ";
    for (unsigned int i = 0; i < m_grid_index.getDimension(); i++) file << " " << m_grid_index.getLength(i);
    file << std::endl;
    file << "#num_gaussians: " << getNumGaussians() << std::endl;
    for (auto& v : m_variables) file << v.m_cv->getName() << m_delimiter;
    file << "grid_value" << m_delimiter << "det_sigma" << m_delimiter << "num_gaussians" << m_delimiter << "hist" << m_delimiter
         << "hist_reweight" << m_delimiter << "weight" << std::endl;
    std::vector<unsigned int> coords(m_grid_index.getDimension());
    for (unsigned int g = 0; g < len; g++) {
        m_grid_index.getCoordinates(g, coords);
        for (size_t i = 0; i < m_variables.size(); ++i) {
            const Scalar delta = (m_variables[i].m_cv_max - m_variables[i].m_cv_min) / (m_variables[i].m_num_points - 1);
            file << std::setprecision(10) << m_variables[i].m_cv_min + coords[i] * delta << m_delimiter;
        }
        file << std::setprecision(10) << grid[g];
        const double val = hist_gauss[g] > 0 ? sigma[g] / (double)hist_gauss[g] : 0.0;
        file << m_delimiter << std::setprecision(10) << val;
        file << m_delimiter << hist_gauss[g] << m_delimiter << hist[g];
        file << m_delimiter << std::setprecision(10) << rew[g];
        file << m_delimiter << std::setprecision(10) << weight[g] << std::endl;
    }
}

// IntegratorMetaDynamics.cc:928-1000
void IntegratorMetaDynamics::readGrid(const std::string& filename) {
    if (!m_use_grid || !m_grid) {
        m_exec_conf->msg->error("integrate.mode_metadynamics: Grid information can only be read if grid is enabled.");
        throw std::runtime_error("Error reading grid.");
    }
    std::ifstream file(filename.c_str());
    std::string line, tmp;
    std::getline(file, line); std::getline(file, line);
    std::getline(file, line);
    unsigned int num_gaussians = 0;
    { std::istringstream iss(line); iss >> tmp >> num_gaussians; }
    std::getline(file, line);
    const unsigned int len = m_grid_index.getNumElements();
    std::vector<double> grid(len), sigma(len), rew(len), weight(len);
    std::vector<unsigned int> hist(len), hist_gauss(len);
    for (unsigned int g = 0; g < len; g++) {
        if (!file.good()) {
            m_exec_conf->msg->error("integrate.mode_metadynamics: Premature end of grid file.");
            throw std::runtime_error("Error reading grid.");
        }
        std::getline(file, line);
        std::istringstream iss(line);
        for (size_t i = 0; i < m_variables.size(); i++) iss >> tmp;
        iss >> grid[g] >> sigma[g] >> hist_gauss[g] >> hist[g];
        sigma[g] *= hist_gauss[g];
        iss >> rew[g] >> weight[g];
    }
    metad_check(metad_grid_upload(m_grid, 0, grid.data()), "grid upload");
    metad_check(metad_grid_upload(m_grid, 1, rew.data()), "grid upload");
    metad_check(metad_grid_upload(m_grid, 2, weight.data()), "grid upload");
    metad_check(metad_grid_upload(m_grid, 3, sigma.data()), "grid upload");
    metad_check(metad_grid_upload(m_grid, 4, hist.data()), "grid upload");
    metad_check(metad_grid_upload(m_grid, 5, hist_gauss.data()), "grid upload");
    metad_check(metad_grid_set_num_gaussians(m_grid, num_gaussians), "grid upload");
}

// IntegratorMetaDynamics.cc:74-119
void IntegratorMetaDynamics::openOutputFile() {
    struct stat buffer;
    const bool file_exists = stat(m_filename.c_str(), &buffer) == 0;
    if (file_exists && !m_overwrite) {
        m_file.open(m_filename.c_str(), std::ios_base::in | std::ios_base::out | std::ios_base::ate);
        m_is_appending = true;
    } else {
        m_file.open(m_filename.c_str(), std::ios_base::out);
        m_is_appending = false;
    }
    if (!m_file.good()) {
        m_exec_conf->msg->error("integrate.mode_metadynamics: Error opening log file " + m_filename);
        throw std::runtime_error("Error initializing IntegratorMetadynamics");
    }
}
void IntegratorMetaDynamics::writeFileHeader() {
    m_file << "timestep" << m_delimiter << "W" << m_delimiter;
    for (size_t i = 0; i < m_variables.size(); ++i) {
        m_file << m_variables[i].m_cv->getName();
        for (size_t j = 0; j < m_variables.size(); ++j)
            m_file << m_delimiter << "sigma_" << m_variables[i].m_cv->getName() << "_" << i << "_" << j;
        m_file << m_delimiter;
    }
    m_file << std::endl;
}

}  // namespace metadynamics
