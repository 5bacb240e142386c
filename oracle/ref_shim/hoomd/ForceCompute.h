// Stand-in for the HOOMD-blue 2.x headers that the reference plugin's CPU sources include (<hoomd/ForceCompute.h> pulls in
// everything they touch).  It exists for ONE purpose: to compile the reference's own .cc files, unmodified and where they
// lie under /root/reference, into oracle/_ref/libref.so, so that the oracle's restatement can be checked against the
// reference's own arithmetic (oracle/Makefile target `_ref`, tests/test_reference_build.py).  TEST INFRASTRUCTURE ONLY:
// nothing in the product path includes or links this.
//
// What is real and what is a stand-in:
//   * the plugin code under test (CollectiveVariable.cc, LamellarOrderParameter.cc, OrderParameterMesh.cc, IndexGrid.cc,
//     AspectRatio.cc, Density.cc) is the reference's, byte for byte;
//   * everything in this file replaces HOOMD.  Containers, handles, messenger, profiler are trivial.  BoxDim is restated
//     from the public HOOMD 2.x header (hoomd/BoxDim.h): lo/hi/L/Linv members with Linv = 1/(hi - lo),
//     makeFraction = ((v - lo) - tilt terms) * Linv (+ ghost fraction, zero here), makeCoordinates = lo + f*L (+ shear),
//     host-side branching minImage, lattice vectors, volume, nearest plane distance.  HOOMD itself is not installable
//     here, so that restatement is the remaining unpinned piece.
#pragma once
#include <cassert>
#include <cmath>
#include <cstring>
#include <functional>
#include <iostream>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>
#include <bitset>
#include <algorithm>

#ifdef SINGLE_PRECISION
typedef float Scalar;
#else
typedef double Scalar;
#endif

struct Scalar2 { Scalar x, y; };
struct Scalar3 { Scalar x, y, z; };
struct Scalar4 { Scalar x, y, z, w; };
struct int3 { int x, y, z; };
struct uint3 { unsigned int x, y, z; };
struct uchar3 { unsigned char x, y, z; };
inline Scalar2 make_scalar2(Scalar x, Scalar y) { Scalar2 r = {x, y}; return r; }
inline Scalar3 make_scalar3(Scalar x, Scalar y, Scalar z) { Scalar3 r = {x, y, z}; return r; }
inline Scalar4 make_scalar4(Scalar x, Scalar y, Scalar z, Scalar w) { Scalar4 r = {x, y, z, w}; return r; }
inline int3 make_int3(int x, int y, int z) { int3 r = {x, y, z}; return r; }
inline uint3 make_uint3(unsigned int x, unsigned int y, unsigned int z) { uint3 r = {x, y, z}; return r; }
inline uchar3 make_uchar3(unsigned char x, unsigned char y, unsigned char z) { uchar3 r = {x, y, z}; return r; }

// VectorMath.h / HOOMDMath.h operators on Scalar3
inline Scalar3 operator+(const Scalar3& a, const Scalar3& b) { return make_scalar3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline Scalar3 operator-(const Scalar3& a, const Scalar3& b) { return make_scalar3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline Scalar3 operator*(const Scalar3& a, const Scalar3& b) { return make_scalar3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline Scalar3 operator/(const Scalar3& a, const Scalar3& b) { return make_scalar3(a.x / b.x, a.y / b.y, a.z / b.z); }
inline Scalar3 operator-(const Scalar3& a) { return make_scalar3(-a.x, -a.y, -a.z); }
inline Scalar3 operator*(const Scalar3& a, const Scalar& b) { return make_scalar3(a.x * b, a.y * b, a.z * b); }
inline Scalar3 operator*(const Scalar& b, const Scalar3& a) { return make_scalar3(a.x * b, a.y * b, a.z * b); }
inline Scalar3 operator/(const Scalar3& a, const Scalar& b) { return make_scalar3(a.x / b, a.y / b, a.z / b); }
inline Scalar3 operator/(const Scalar& b, const Scalar3& a) { return make_scalar3(b / a.x, b / a.y, b / a.z); }
inline Scalar3& operator+=(Scalar3& a, const Scalar3& b) { a.x += b.x; a.y += b.y; a.z += b.z; return a; }
inline Scalar3& operator-=(Scalar3& a, const Scalar3& b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; return a; }
inline Scalar3& operator*=(Scalar3& a, const Scalar& b) { a.x *= b; a.y *= b; a.z *= b; return a; }
inline Scalar3& operator/=(Scalar3& a, const Scalar& b) { a.x /= b; a.y /= b; a.z /= b; return a; }
inline Scalar dot(const Scalar3& a, const Scalar3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Scalar3 cross(const Scalar3& a, const Scalar3& b) {
    return make_scalar3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
namespace fast {
inline Scalar sqrt(Scalar x) { return ::sqrt(x); }
inline Scalar rsqrt(Scalar x) { return Scalar(1.0) / ::sqrt(x); }
inline Scalar sin(Scalar x) { return ::sin(x); }
inline Scalar cos(Scalar x) { return ::cos(x); }
inline Scalar exp(Scalar x) { return ::exp(x); }
inline Scalar pow(Scalar x, Scalar y) { return ::pow(x, y); }
}  // namespace fast
namespace slow {
inline Scalar rint(Scalar x) { return ::rint(x); }
inline Scalar floor(Scalar x) { return ::floor(x); }
inline Scalar sqrt(Scalar x) { return ::sqrt(x); }
}  // namespace slow
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#define HOSTDEVICE inline
#define TAG_ALLOCATION(x)
// HOOMDMath.h: the particle type lives in the low 32 bits of the w component
inline int __scalar_as_int(Scalar b) { int i; memcpy(&i, &b, sizeof i); return i; }
inline Scalar __int_as_scalar(int a) { Scalar b = 0; memcpy(&b, &a, sizeof a); return b; }

// ---- BoxDim (hoomd/BoxDim.h, restated) -------------------------------------------------------------------
struct BoxDim {
    Scalar3 m_lo, m_hi, m_L, m_Linv;
    Scalar m_xy, m_xz, m_yz;
    uchar3 m_periodic;
    BoxDim() { init(1, 1, 1); }
    explicit BoxDim(Scalar Len) { init(Len, Len, Len); }
    BoxDim(Scalar Lx, Scalar Ly, Scalar Lz) { init(Lx, Ly, Lz); }
    void init(Scalar Lx, Scalar Ly, Scalar Lz) {
        setLoHi(make_scalar3(-Lx / Scalar(2.0), -Ly / Scalar(2.0), -Lz / Scalar(2.0)),
                make_scalar3(Lx / Scalar(2.0), Ly / Scalar(2.0), Lz / Scalar(2.0)));
        m_xy = m_xz = m_yz = Scalar(0.0);
        m_periodic = make_uchar3(1, 1, 1);
    }
    void setLoHi(const Scalar3& lo, const Scalar3& hi) {
        m_lo = lo; m_hi = hi;
        m_Linv = Scalar(1.0) / (m_hi - m_lo);
        m_L = m_hi - m_lo;
    }
    void setTiltFactors(Scalar xy, Scalar xz, Scalar yz) { m_xy = xy; m_xz = xz; m_yz = yz; }
    Scalar getTiltFactorXY() const { return m_xy; }
    Scalar getTiltFactorXZ() const { return m_xz; }
    Scalar getTiltFactorYZ() const { return m_yz; }
    uchar3 getPeriodic() const { return m_periodic; }
    Scalar3 getL() const { return m_L; }
    Scalar3 getLo() const { return m_lo; }
    Scalar3 getHi() const { return m_hi; }
    Scalar3 makeFraction(const Scalar3& v, const Scalar3& ghost_width = make_scalar3(0.0, 0.0, 0.0)) const {
        Scalar3 ghost_frac = ghost_width / getNearestPlaneDistance();
        Scalar3 delta = v - m_lo;
        delta.x -= (m_xz - m_yz * m_xy) * v.z + m_xy * v.y;
        delta.y -= m_yz * v.z;
        return (delta * m_Linv + ghost_frac) / (make_scalar3(1, 1, 1) + Scalar(2.0) * ghost_frac);
    }
    Scalar3 makeCoordinates(const Scalar3& f) const {
        Scalar3 v = m_lo + f * m_L;
        v.x += m_xy * v.y + m_xz * v.z;
        v.y += m_yz * v.z;
        return v;
    }
    Scalar3 minImage(const Scalar3& v) const {       // host variant: branches, one box length per direction
        Scalar3 w = v;
        Scalar3 L = getL();
        if (m_periodic.z) {
            if (w.z >= m_hi.z) { w.z -= L.z; w.y -= L.z * m_yz; w.x -= L.z * m_xz; }
            else if (w.z < m_lo.z) { w.z += L.z; w.y += L.z * m_yz; w.x += L.z * m_xz; }
        }
        if (m_periodic.y) {
            if (w.y >= m_hi.y) { w.y -= L.y; w.x -= L.y * m_xy; }
            else if (w.y < m_lo.y) { w.y += L.y; w.x += L.y * m_xy; }
        }
        if (m_periodic.x) {
            if (w.x >= m_hi.x) w.x -= L.x;
            else if (w.x < m_lo.x) w.x += L.x;
        }
        return w;
    }
    Scalar3 getLatticeVector(unsigned int i) const {
        if (i == 0) return make_scalar3(m_L.x, 0.0, 0.0);
        if (i == 1) return make_scalar3(m_L.y * m_xy, m_L.y, 0.0);
        if (i == 2) return make_scalar3(m_L.z * m_xz, m_L.z * m_yz, m_L.z);
        return make_scalar3(0.0, 0.0, 0.0);
    }
    Scalar getVolume(bool twod = false) const { return twod ? m_L.x * m_L.y : m_L.x * m_L.y * m_L.z; }
    Scalar3 getNearestPlaneDistance() const {
        Scalar3 dist;
        dist.x = m_L.x / ::sqrt(Scalar(1.0) + m_xy * m_xy + (m_xy * m_yz - m_xz) * (m_xy * m_yz - m_xz));
        dist.y = m_L.y / ::sqrt(Scalar(1.0) + m_yz * m_yz);
        dist.z = m_L.z;
        return dist;
    }
};

// ---- trivial containers / handles ---------------------------------------------------------------------------
struct access_location { enum Enum { host, device }; };
struct access_mode { enum Enum { read, readwrite, overwrite }; };
class ExecutionConfiguration;
template <class T> class GPUArray {
  public:
    GPUArray() : m_pitch(0), m_height(1) {}
    GPUArray(unsigned int n, std::shared_ptr<const ExecutionConfiguration>) : m_data(n), m_pitch(n), m_height(1) {}
    GPUArray(unsigned int w, unsigned int h, std::shared_ptr<const ExecutionConfiguration>)
        : m_data((size_t)(w + ((16 - (w & 15)) & 15)) * h), m_pitch(w + ((16 - (w & 15)) & 15)), m_height(h) {}
    void swap(GPUArray& o) { m_data.swap(o.m_data); std::swap(m_pitch, o.m_pitch); std::swap(m_height, o.m_height); }
    unsigned int getNumElements() const { return (unsigned int)m_data.size(); }
    unsigned int getPitch() const { return m_pitch; }
    unsigned int getHeight() const { return m_height; }
    bool isNull() const { return m_data.empty(); }
    void resize(unsigned int n) { m_data.resize(n); m_pitch = n; }
    mutable std::vector<T> m_data;
    unsigned int m_pitch, m_height;
};
template <class T> class GlobalArray : public GPUArray<T> {
  public:
    GlobalArray() {}
    GlobalArray(unsigned int n, std::shared_ptr<const ExecutionConfiguration> e, const std::string& = std::string()) : GPUArray<T>(n, e) {}
    GlobalArray(unsigned int w, unsigned int h, std::shared_ptr<const ExecutionConfiguration> e) : GPUArray<T>(w, h, e) {}
};
template <class T> using GlobalVector = GlobalArray<T>;
template <class T> class ArrayHandle {
  public:
    ArrayHandle(const GPUArray<T>& a, access_location::Enum = access_location::host, access_mode::Enum = access_mode::readwrite)
        : data(a.m_data.data()) {}
    T* const data;
};

// ---- messenger, profiler, execution configuration ------------------------------------------------------------
class Messenger {
  public:
    std::ostream& error() const { return m_sink; }
    std::ostream& warning() const { return m_sink; }
    std::ostream& notice(unsigned int) const { return m_sink; }
    mutable std::ostringstream m_sink;
};
class Profiler {
  public:
    void push(const std::string&) {}
    template <class E> void push(E, const std::string&) {}
    void pop() {}
    template <class E> void pop(E) {}
    template <class E, class A, class B> void pop(E, A, B) {}
};
struct cudaDeviceProp { int warpSize = 32, maxThreadsPerBlock = 1024; };      // named by WellTemperedEnsemble.cc's GPU branch
class GPUPartition {};
struct CachedAllocator {};
class ExecutionConfiguration {
  public:
    ExecutionConfiguration() : msg(new Messenger()) {}
    enum executionMode { GPU, CPU, AUTO };
    executionMode exec_mode = CPU;
    cudaDeviceProp dev_prop;
    void beginMultiGPU() const {}
    void endMultiGPU() const {}
    unsigned int getNumActiveGPUs() const { return 1; }
    const CachedAllocator& getCachedAllocatorManaged() const { static CachedAllocator a; return a; }
    std::shared_ptr<Messenger> msg;
    bool isCUDAEnabled() const { return false; }
    bool isCUDAErrorCheckingEnabled() const { return false; }
    unsigned int getRank() const { return 0; }
    unsigned int getNRanks() const { return 1; }
    bool isRoot() const { return true; }
    unsigned int getPartition() const { return 0; }
};

// ---- particle data, system definition -------------------------------------------------------------------------
struct pdata_flag { enum Enum { isotropic_virial = 0, potential_energy, pressure_tensor, rotational_kinetic_energy, external_field_virial }; };
typedef std::bitset<32> PDataFlags;
namespace Nano { template <class Sig> class Signal {
  public:
    template <class T, void (T::*M)()> void connect(T*) {}
    template <class T, void (T::*M)()> void disconnect(T*) {}
}; }
class DomainDecomposition;
class ParticleData {
  public:
    ParticleData(unsigned int N, const BoxDim& box, unsigned int ntypes, std::shared_ptr<ExecutionConfiguration> exec)
        : m_box(box), m_N(N), m_ntypes(ntypes), m_exec_conf(exec), m_pos(N, exec), m_net_force(N, exec), m_net_torque(N, exec),
          m_net_virial(N, 6, exec), m_external_energy(0) { for (int i = 0; i < 6; ++i) m_external_virial[i] = 0; }
    const BoxDim& getBox() const { return m_box; }
    const BoxDim& getGlobalBox() const { return m_box; }
    void setGlobalBox(const BoxDim& b) { m_box = b; }
    unsigned int getN() const { return m_N; }
    unsigned int getNGlobal() const { return m_N; }
    unsigned int getMaxN() const { return m_N; }
    unsigned int getNGhosts() const { return 0; }
    unsigned int getNTypes() const { return m_ntypes; }
    const GlobalArray<Scalar4>& getPositions() const { return m_pos; }
    const GlobalArray<Scalar4>& getNetForce() const { return m_net_force; }
    const GlobalArray<Scalar4>& getNetTorqueArray() const { return m_net_torque; }
    const GlobalArray<Scalar>& getNetVirial() const { return m_net_virial; }
    Scalar getExternalEnergy() const { return m_external_energy; }
    Scalar getExternalVirial(unsigned int i) const { return m_external_virial[i]; }
    void setExternalVirial(unsigned int i, Scalar v) { m_external_virial[i] = v; }
    void setExternalEnergy(Scalar e) { m_external_energy = e; }
    PDataFlags getFlags() const { return m_flags; }
    void setFlags(const PDataFlags& f) { m_flags = f; }
    std::shared_ptr<DomainDecomposition> getDomainDecomposition() const { return std::shared_ptr<DomainDecomposition>(); }
    const GPUPartition& getGPUPartition() const { static GPUPartition g; return g; }
    Nano::Signal<void()>& getBoxChangeSignal() { return m_box_signal; }
    std::shared_ptr<ExecutionConfiguration> getExecConf() const { return m_exec_conf; }
    BoxDim m_box;
    unsigned int m_N, m_ntypes;
    std::shared_ptr<ExecutionConfiguration> m_exec_conf;
    GlobalArray<Scalar4> m_pos, m_net_force, m_net_torque;
    GlobalArray<Scalar> m_net_virial;
    Scalar m_external_energy, m_external_virial[6];
    PDataFlags m_flags;
    Nano::Signal<void()> m_box_signal;
};
class SystemDefinition {
  public:
    explicit SystemDefinition(std::shared_ptr<ParticleData> p) : m_pdata(p) {}
    std::shared_ptr<ParticleData> getParticleData() const { return m_pdata; }
    unsigned int getNDimensions() const { return 3; }
    std::shared_ptr<ParticleData> m_pdata;
};
class Communicator;

// ---- ForceCompute (hoomd/ForceCompute.h): the members the plugin's classes use ----------------------------------
class ForceCompute {
  public:
    explicit ForceCompute(std::shared_ptr<SystemDefinition> sysdef)
        : m_sysdef(sysdef), m_pdata(sysdef->getParticleData()), m_exec_conf(m_pdata->getExecConf()), m_prof(),
          m_force(m_pdata->getN(), m_exec_conf), m_virial(m_pdata->getN(), 6, m_exec_conf), m_torque(m_pdata->getN(), m_exec_conf),
          m_external_energy(0) {
        m_virial_pitch = m_virial.getPitch();
        for (int i = 0; i < 6; ++i) m_external_virial[i] = 0;
    }
    virtual ~ForceCompute() {}
    virtual std::vector<std::string> getProvidedLogQuantities() { return std::vector<std::string>(); }
    virtual Scalar getLogValue(const std::string&, unsigned int) { return Scalar(0.0); }
    virtual void compute(unsigned int timestep) {
        // ForceCompute::compute zeroes nothing itself; the plugin's classes overwrite every entry they own
        computeForces(timestep);
    }
    GlobalArray<Scalar4>& getForceArray() { return m_force; }
    GlobalArray<Scalar>& getVirialArray() { return m_virial; }
    Scalar getExternalVirial(unsigned int i) const { return m_external_virial[i]; }
    Scalar getExternalEnergy() const { return m_external_energy; }
    virtual void setAutotunerParams(bool, unsigned int) {}
  protected:
    virtual void computeForces(unsigned int timestep) = 0;
    std::shared_ptr<SystemDefinition> m_sysdef;
    std::shared_ptr<ParticleData> m_pdata;
    std::shared_ptr<const ExecutionConfiguration> m_exec_conf;
    std::shared_ptr<Profiler> m_prof;
    std::shared_ptr<Communicator> m_comm;
    GlobalArray<Scalar4> m_force;
    GlobalArray<Scalar> m_virial;
    GlobalArray<Scalar4> m_torque;
    unsigned int m_virial_pitch;
    Scalar m_external_virial[6];
    Scalar m_external_energy;
};
