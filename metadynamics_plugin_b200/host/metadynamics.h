// metadynamics.h -- host classes with the reference plugin's operator surface, implemented over the C ABI
// (include/metad_b200.h).  Same class names, virtuals, setters and error behaviour as the reference headers:
//   CollectiveVariable            CollectiveVariable.h:32-196, CollectiveVariable.cc:22-106
//   LamellarOrderParameterGPU     LamellarOrderParameterGPU.h/.cc (LamellarOrderParameter.h:32-101)
//   OrderParameterMeshGPU         OrderParameterMeshGPU.h:36-56 (OrderParameterMesh.h)
//   WellTemperedEnsemble          WellTemperedEnsemble.h/.cc
//   AspectRatio, Density          AspectRatio.cc, Density.cc (host scalars)
//   IndexGrid                     IndexGrid.h/.cc
//   IntegratorMetaDynamics        IntegratorMetaDynamics.h:69-383, .cc:121-588, 778-1000
// Differences that are deliberate: the CV value and the bias factor live in device memory (double), so a step
// is enqueued without host round trips; getCurrentValue() remains the synchronising host accessor.
#pragma once
#include <fstream>
#include <functional>
#include <map>

#include "hoomd_shim.h"

namespace metadynamics {
using namespace shim;

class CollectiveVariable : public ForceCompute {
  public:
    enum umbrella_Enum { no_umbrella = 0, linear, harmonic, wall, gaussian };

    CollectiveVariable(std::shared_ptr<SystemDefinition> sysdef, const std::string& name);
    virtual ~CollectiveVariable() {}

    //! Current value on the host (synchronises with the device)
    virtual Scalar getCurrentValue(unsigned int timestep) { return Scalar(0.0); }
    //! Device copy of the current value (double); default implementation uploads getCurrentValue()
    virtual const double* getCurrentValueDevice(unsigned int timestep);
    //! Set the bias factor dV/ds from the host
    virtual void setBiasFactor(Scalar bias);
    //! Hand over the bias factor from device memory (no host round trip)
    void setBiasFactorDevice(const double* d_bias);

    void setUmbrella(umbrella_Enum umbrella) { m_umbrella = umbrella; if (umbrella == no_umbrella) setBiasFactor(Scalar(0.0)); }
    void setKappa(Scalar kappa) { m_kappa = kappa; }
    void setWidthFlat(Scalar width) { m_width_flat = width; }
    void setScale(Scalar scale) { m_scale = scale; }
    void setMinimum(Scalar cv0) { m_cv0 = cv0; }
    std::string getName() { return m_cv_name; }
    void computeDerivatives(unsigned int timestep) { setBiasFactor(Scalar(1.0)); computeBiasForces(timestep); }
    virtual bool canComputeDerivatives() { return true; }
    Scalar getUmbrellaPotential(unsigned int timestep);
    virtual bool requiresNetForce() { return false; }
    std::vector<std::string> getProvidedLogQuantities() override { return {"umbrella_energy_" + m_cv_name}; }
    Scalar getLogValue(const std::string& quantity, unsigned int timestep) override;
    //! host mirror of the bias factor (synchronises)
    Scalar getBiasFactor();

  protected:
    void computeForces(unsigned int timestep) override;
    virtual void computeBiasForces(unsigned int timestep) {}
    const double* biasDevice() const { return m_d_scalars.data() + (m_bias_with_umbrella ? 1 : 0); }
    // host copy of the factor that computeBiasForces must apply: the integrator's dV/ds PLUS the umbrella increment of
    // computeForces (the reference's m_bias at the time it calls computeBiasForces, CollectiveVariable.cc:22-66)
    Scalar biasHost();

    std::string m_cv_name;
    DeviceArray<double> m_d_scalars;   // [0] bias factor from the integrator/host, [1] bias incl. umbrella, [2] CV value,
                                       // [3] the factor of the last computeBiasForces (kept for a lazily evaluated external virial)
    // host-scalar CVs: remember the factor of this computeBiasForces on the device; the external virial is evaluated from it
    // when somebody reads it (updateExternalVirial), so the step itself has no device -> host round trip
    void keepBiasForVirial();
    Scalar keptBias();
    bool m_virial_dirty = false;
    bool m_bias_with_umbrella = false;

  private:
    umbrella_Enum m_umbrella = no_umbrella;
    Scalar m_cv0 = 0, m_kappa = 1, m_width_flat = 0, m_scale = 1;
};

class LamellarOrderParameterGPU : public CollectiveVariable {
  public:
    LamellarOrderParameterGPU(std::shared_ptr<SystemDefinition> sysdef, const std::vector<Scalar>& mode,
                              const std::vector<int3_>& lattice_vectors, const std::string& suffix = "");
    ~LamellarOrderParameterGPU();
    Scalar getCurrentValue(unsigned int timestep) override;            // always recomputes (LamellarOrderParameter.h:75-79)
    const double* getCurrentValueDevice(unsigned int timestep) override;
    std::vector<std::string> getProvidedLogQuantities() override;
    Scalar getLogValue(const std::string& quantity, unsigned int timestep) override;
    //! hook for particle-sharded runs: all-reduce (sum) of 2*n_wave doubles in device memory between modes and CV
    std::function<void(double* d_modes, int n)> allreduce;

  protected:
    void computeBiasForces(unsigned int timestep) override;
    virtual void computeCV(unsigned int timestep);

    std::string m_log_name;
    metad_lamellar* m_plan = nullptr;
    int m_n_wave = 0;
    DeviceArray<double> m_d_modes;
    unsigned int m_cv_last_updated = 0;
};

class OrderParameterMeshGPU : public CollectiveVariable {
  public:
    OrderParameterMeshGPU(std::shared_ptr<SystemDefinition> sysdef, unsigned int nx, unsigned int ny, unsigned int nz,
                          std::vector<Scalar> mode, std::vector<int3_> zero_modes);
    ~OrderParameterMeshGPU();
    Scalar getCurrentValue(unsigned int timestep) override;            // cached per timestep (OrderParameterMesh.cc:927-928)
    const double* getCurrentValueDevice(unsigned int timestep) override;
    std::vector<std::string> getProvidedLogQuantities() override;
    Scalar getLogValue(const std::string& quantity, unsigned int timestep) override;
    void setTable(const std::vector<Scalar>& K, const std::vector<Scalar>& d_K, Scalar kmin, Scalar kmax);
    void setUseTable(bool use_table);

  protected:
    void computeBiasForces(unsigned int timestep) override;

    void computeQmax(unsigned int timestep);       // OrderParameterMesh.cc:1108-1179
    void computeVirial();                          // OrderParameterMesh.cc:970-1050
    void enableExtras();

    metad_mesh* m_plan = nullptr;
    bool m_is_first_step = true, m_use_table = false, m_extras = false;
    unsigned int m_cv_last_updated = 0, m_q_max_last_computed = 0;
    Scalar m_q_max[3] = {0, 0, 0}, m_sq_max = 0;
    std::vector<Scalar> m_table, m_table_d;
    Scalar m_k_min = 0, m_k_max = 0;
};

class WellTemperedEnsemble : public CollectiveVariable {
  public:
    WellTemperedEnsemble(std::shared_ptr<SystemDefinition> sysdef, const std::string& name);
    bool requiresNetForce() override { return true; }
    bool canComputeDerivatives() override { return false; }
    Scalar getCurrentValue(unsigned int timestep) override;
    const double* getCurrentValueDevice(unsigned int timestep) override;
    std::vector<std::string> getProvidedLogQuantities() override;
    Scalar getLogValue(const std::string& quantity, unsigned int timestep) override;

  protected:
    void computeBiasForces(unsigned int timestep) override;
    std::string m_log_name;
};

class AspectRatio : public CollectiveVariable {
  public:
    AspectRatio(std::shared_ptr<SystemDefinition> sysdef, unsigned int dir1, unsigned int dir2);
    Scalar getCurrentValue(unsigned int timestep) override;
    bool canComputeDerivatives() override { return false; }
    void updateExternalVirial() override;

  protected:
    void computeBiasForces(unsigned int timestep) override;
    unsigned int m_dir1, m_dir2;
};

class Density : public CollectiveVariable {
  public:
    Density(std::shared_ptr<SystemDefinition> sysdef, const std::string& suffix);
    Scalar getCurrentValue(unsigned int timestep) override;
    bool canComputeDerivatives() override { return false; }
    void updateExternalVirial() override;

  protected:
    void computeBiasForces(unsigned int timestep) override;
};

//! Wraps a CollectiveVariable around a regular ForceCompute: CV = its potential energy (CollectiveWrapper.cc:31-73), the
//! bias force = its own force, torque and virial arrays scaled IN PLACE by the bias factor (:140-188 -- by m_bias, not
//! 1 + m_bias as in WellTemperedEnsemble; restated as written).  Reuses the WTE kernels, like the reference.
class CollectiveWrapper : public CollectiveVariable {
  public:
    CollectiveWrapper(std::shared_ptr<SystemDefinition> sysdef, std::shared_ptr<ForceCompute> fc, const std::string& name);
    Scalar getCurrentValue(unsigned int timestep) override;
    const double* getCurrentValueDevice(unsigned int timestep) override;
    //! sum over the ranks of one walker of the local energy (MPI_Allreduce, CollectiveWrapper.cc:63-69): device double, in place
    std::function<void(size_t, size_t)> allreduce;

  protected:
    void computeBiasForces(unsigned int timestep) override;
    std::shared_ptr<ForceCompute> m_fc;
    DeviceArray<double> m_d_fac;
};

class IndexGrid {
  public:
    IndexGrid();
    explicit IndexGrid(const std::vector<unsigned int>& lengths) { setLengths(lengths); }
    void setLengths(const std::vector<unsigned int>& lengths);
    unsigned int getIndex(const std::vector<unsigned int>& coords);
    void getCoordinates(const unsigned int idx, std::vector<unsigned int>& coords);
    unsigned int getNumElements();
    unsigned int getLength(const unsigned int i) { return m_lengths.at(i); }
    unsigned int getDimension() { return (unsigned int)m_lengths.size(); }

  private:
    std::vector<unsigned int> m_lengths, m_factors;
};

struct CollectiveVariableItem {
    std::shared_ptr<CollectiveVariable> m_cv;
    Scalar m_sigma, m_cv_min, m_cv_max;
    unsigned int m_num_points;
};

//! Stand-in for HOOMD's System: owns the list of force computes the integrator sums into the net force
class System {
  public:
    explicit System(std::shared_ptr<SystemDefinition> sysdef) : m_sysdef(sysdef) {}
    void addCompute(std::shared_ptr<ForceCompute> fc, const std::string& name) { m_computes.push_back(fc); m_names.push_back(name); }
    const std::vector<std::shared_ptr<ForceCompute>>& computes() const { return m_computes; }
    std::shared_ptr<SystemDefinition> getSystemDefinition() const { return m_sysdef; }

  private:
    std::shared_ptr<SystemDefinition> m_sysdef;
    std::vector<std::shared_ptr<ForceCompute>> m_computes;
    std::vector<std::string> m_names;
};

class IntegratorMetaDynamics {
  public:
    enum Enum { mode_standard, mode_well_tempered };

    IntegratorMetaDynamics(std::shared_ptr<SystemDefinition> sysdef, Scalar deltaT, Scalar W, Scalar T_shift, Scalar T,
                           unsigned int stride, bool add_bias = true, const std::string& filename = "", bool overwrite = false,
                           const Enum mode = mode_standard);
    ~IntegratorMetaDynamics();

    void setSystem(std::shared_ptr<System> system) { m_system = system; }
    void update(unsigned int timestep);
    void prepRun(unsigned int timestep);
    void registerCollectiveVariable(std::shared_ptr<CollectiveVariable> cv, Scalar sigma, Scalar cv_min = Scalar(0.0),
                                    Scalar cv_max = Scalar(0.0), int num_points = 0);
    void removeAllVariables() { m_variables.clear(); }
    std::vector<std::string> getProvidedLogQuantities() { return m_log_names; }
    Scalar getLogValue(const std::string& quantity, unsigned int timestep);
    void setGrid(bool use_grid);
    void setMode(Enum mode);
    void setStride(unsigned int stride);
    bool isInitialized() { return m_is_initialized; }
    void dumpGrid(const std::string& filename1, const std::string& filename2, unsigned int period);
    void restartFromGridFile(const std::string& filename) { m_restart_filename = filename; }
    void setAddHills(bool add_bias);
    void setAdaptive(bool adaptive);
    void setSigmaG(Scalar sigma_g) { m_sigma_g = sigma_g; }
    void setMultipleWalkers(bool multiple) { m_multiple_walkers = multiple; }
    //! Multiple walkers: sum two device buffers in place over the walkers (the partition communicator of the reference,
    //! IntegratorMetaDynamics.cc:65-71): called with (device pointer of n_d doubles, n_d, device pointer of n_u unsigned, n_u)
    //! on deposit steps; must have completed when it returns.  Unset: one walker (nothing to share).
    std::function<void(size_t, size_t, size_t, size_t)> walker_allreduce;
    //! Domain decomposition: sum n doubles in device memory over the ranks of one walker (computeSigma's MPI_Allreduce, :1265-1274)
    std::function<void(size_t, size_t)> domain_allreduce;
    //! the inverse sigma matrix in use (adaptive Gaussians; row-major n_cv x n_cv)
    std::vector<double> getSigmaInv() const { return m_sigma_inv; }
    void resetHistogram();
    //! bias grid arrays for inspection: "grid", "reweighted", "weight", "sigma_grid" (double), "hist", "hist_gauss" (as double)
    std::vector<double> getGridArray(const std::string& name);
    unsigned int getNumGaussians();

  private:
    void updateBiasPotential(unsigned int timestep);
    void computeSigma();
    void computeNetForce(unsigned int timestep);
    void setupGrid();
    void readGrid(const std::string& filename);
    void writeGrid(const std::string& filename, unsigned int timestep);
    void openOutputFile();
    void writeFileHeader();
    void pushFlags();

    std::shared_ptr<SystemDefinition> m_sysdef;
    std::shared_ptr<ParticleData> m_pdata;
    std::shared_ptr<ExecutionConfiguration> m_exec_conf;
    std::shared_ptr<System> m_system;
    Scalar m_deltaT, m_W, m_T_shift;
    unsigned int m_stride;
    std::vector<CollectiveVariableItem> m_variables;
    std::vector<std::string> m_log_names;
    bool m_is_initialized = false, m_prepared = false;
    const std::string m_filename;
    bool m_overwrite, m_is_appending = false;
    std::ofstream m_file;
    std::string m_delimiter = "\t";
    bool m_use_grid = false;
    IndexGrid m_grid_index;
    bool m_add_bias;
    std::string m_restart_filename, m_grid_fname1, m_grid_fname2;
    unsigned int m_grid_period = 0, m_cur_file = 0;
    Scalar m_sigma_g = 1.0;
    bool m_adaptive = false;
    Scalar m_temp;
    Enum m_mode;
    bool m_multiple_walkers = false;
    metad_grid* m_grid = nullptr;
    DeviceArray<double> m_d_cv, m_d_bias;     // n_cv each
    DeviceArray<double> m_d_sigmasq;          // n_cv^2 sums of products of the CV derivatives (adaptive Gaussians)
    DeviceArray<double> m_d_walk_d;           // multiple walkers: grid_delta | sigma_grid_delta
    DeviceArray<unsigned int> m_d_walk_u;     //                   hist_delta | hist_gauss_delta
    std::vector<double> m_sigma_inv;
    double* m_h_pinned = nullptr;             // pinned staging for host-scalar CVs
};

}  // namespace metadynamics
