// Stand-in for hoomd/md/IntegratorTwoStep.h (see ref_shim/hoomd/ForceCompute.h): just enough of HOOMD's integrator base
// for the reference's IntegratorMetaDynamics.cc to compile and run its bias-potential update.  No integration methods, no
// net-force summation (the oracle cross-check only drives updateBiasPotential with prescribed CV values).
// TEST INFRASTRUCTURE ONLY.
#pragma once
#include <hoomd/ForceCompute.h>
#include <fstream>

class Communicator {
  public:
    void communicate(unsigned int) {}
};
class IntegrationMethodTwoStep {
  public:
    virtual ~IntegrationMethodTwoStep() {}
    virtual void integrateStepOne(unsigned int) {}
    virtual void integrateStepTwo(unsigned int) {}
};
class Integrator {
  public:
    Integrator(std::shared_ptr<SystemDefinition> sysdef, Scalar deltaT)
        : m_sysdef(sysdef), m_pdata(sysdef->getParticleData()), m_exec_conf(m_pdata->getExecConf()), m_deltaT(deltaT) {}
    virtual ~Integrator() {}
    virtual std::vector<std::string> getProvidedLogQuantities() { return std::vector<std::string>(); }
    virtual Scalar getLogValue(const std::string&, unsigned int) { return Scalar(0.0); }
  protected:
    std::shared_ptr<SystemDefinition> m_sysdef;
    std::shared_ptr<ParticleData> m_pdata;
    std::shared_ptr<const ExecutionConfiguration> m_exec_conf;
    std::shared_ptr<Profiler> m_prof;
    std::shared_ptr<Communicator> m_comm;
    std::vector<std::shared_ptr<ForceCompute> > m_forces;
    Scalar m_deltaT;
};
class IntegratorTwoStep : public Integrator {
  public:
    IntegratorTwoStep(std::shared_ptr<SystemDefinition> sysdef, Scalar deltaT)
        : Integrator(sysdef, deltaT), m_prepared(false), m_gave_warning(false) {}
    virtual ~IntegratorTwoStep() {}
    virtual void update(unsigned int) {}
    virtual void prepRun(unsigned int) { m_prepared = true; }
  protected:
    void computeNetForce(unsigned int) {}
    void computeNetForceGPU(unsigned int) {}
    void updateRigidBodies(unsigned int) {}
    std::vector<std::shared_ptr<IntegrationMethodTwoStep> > m_methods;
    bool m_prepared, m_gave_warning;
};
