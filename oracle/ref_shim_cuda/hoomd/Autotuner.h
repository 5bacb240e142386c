// Stand-in for hoomd/Autotuner.h: the reference's WellTemperedEnsemble.h names its autotuners outside ENABLE_CUDA guards, so
// its translation unit is compiled with ENABLE_CUDA defined and these inert declarations (the CPU branch is the one that
// runs: ExecutionConfiguration::exec_mode is CPU).  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <memory>
#include <string>
#include <hoomd/ForceCompute.h>
class Autotuner {
  public:
    Autotuner(unsigned int, unsigned int, unsigned int, unsigned int, unsigned int, const std::string&,
              std::shared_ptr<const ExecutionConfiguration>) {}
    void begin() {}
    void end() {}
    unsigned int getParam() const { return 32; }
    void setPeriod(unsigned int) {}
    void setEnabled(bool) {}
};
