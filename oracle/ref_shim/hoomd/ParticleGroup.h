// Stand-in for HOOMD-blue 2.x hoomd/ParticleGroup.h: the one member the reference's Density CV reads (Density.cc:25,41).
// Test infrastructure only (see ForceCompute.h in this directory).
#pragma once
#include <hoomd/ForceCompute.h>
class ParticleGroup {
  public:
    explicit ParticleGroup(unsigned int n_global) : m_n(n_global) {}
    unsigned int getNumMembersGlobal() const { return m_n; }
    unsigned int getNumMembers() const { return m_n; }
  private:
    unsigned int m_n;
};
