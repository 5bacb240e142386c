// module.cc -- pybind11 module `_metadynamics`: the reference's Python-visible classes (module.cc:24-41 and the
// export_* functions at the bottom of every reference .cc) under the same names, plus the HOOMD stand-ins the
// Python layer needs when `hoomd` itself is absent (SystemDefinition, ParticleData, System, BoxDim).
#include <pybind11/functional.h>
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>
#include <pybind11/stl_bind.h>

#include "metadynamics.h"

namespace py = pybind11;
using namespace metadynamics;

PYBIND11_MAKE_OPAQUE(std::vector<shim::int3_>);
PYBIND11_MAKE_OPAQUE(std::vector<shim::Scalar>);

namespace {
typedef py::array_t<float, py::array::c_style | py::array::forcecast> farray;

void upload4(DeviceArray<Scalar4>& dst, const farray& a) {
    if (a.ndim() != 2 || a.shape(1) != 4 || (size_t)a.shape(0) != dst.size()) throw std::runtime_error("expected a float32 array of shape (N,4)");
    if (dst.size()) dst.upload(reinterpret_cast<const Scalar4*>(a.data()), dst.size());
}
farray download4(const DeviceArray<Scalar4>& src) {
    farray out({(py::ssize_t)src.size(), (py::ssize_t)4});
    cuda_check(cudaDeviceSynchronize(), "sync");
    if (src.size()) src.download(reinterpret_cast<Scalar4*>(out.mutable_data()), src.size());
    return out;
}
}  // namespace

PYBIND11_MODULE(_metadynamics, m) {
    m.doc() = "B200-native metadynamics plugin: reference operator surface over the sm_100a C ABI";

    py::bind_vector<std::vector<int3_>>(m, "std_vector_int3");
    py::bind_vector<std::vector<Scalar>>(m, "std_vector_scalar");
    py::class_<int3_>(m, "int3").def(py::init<>()).def_readwrite("x", &int3_::x).def_readwrite("y", &int3_::y).def_readwrite("z", &int3_::z);
    m.def("make_int3", [](int x, int y, int z) { return int3_{x, y, z}; });

    // ---- HOOMD stand-ins ------------------------------------------------------------------------------------
    py::class_<BoxDim>(m, "BoxDim")
        .def(py::init<double, double, double, double, double, double>(), py::arg("Lx"), py::arg("Ly"), py::arg("Lz"),
             py::arg("xy") = 0.0, py::arg("xz") = 0.0, py::arg("yz") = 0.0)
        .def("getL", [](const BoxDim& b) { return py::make_tuple(b.L(0), b.L(1), b.L(2)); })
        .def("getVolume", &BoxDim::getVolume);
    py::class_<Messenger, std::shared_ptr<Messenger>>(m, "Messenger")
        .def_readwrite("notice_level", &Messenger::notice_level)
        .def_readonly("n_warnings", &Messenger::n_warnings)
        .def_readonly("n_errors", &Messenger::n_errors)
        .def_readonly("last_error", &Messenger::last_error);
    py::class_<ExecutionConfiguration, std::shared_ptr<ExecutionConfiguration>>(m, "ExecutionConfiguration")
        .def("isCUDAEnabled", &ExecutionConfiguration::isCUDAEnabled)
        .def_readonly("msg", &ExecutionConfiguration::msg);
    py::class_<ParticleData, std::shared_ptr<ParticleData>>(m, "ParticleData")
        .def("getN", &ParticleData::getN)
        .def("getNGlobal", &ParticleData::getNGlobal)
        .def("setNGlobal", &ParticleData::setNGlobal)
        .def("getNTypes", &ParticleData::getNTypes)
        .def("getNameByType", &ParticleData::getNameByType)
        .def("getGlobalBox", &ParticleData::getGlobalBox)
        .def("setGlobalBox", &ParticleData::setGlobalBox)
        .def("setPositions", [](ParticleData& p, const farray& a) { upload4(p.getPositions(), a); })
        .def("getPositions", [](ParticleData& p) { return download4(p.getPositions()); })
        .def("setNetForce", [](ParticleData& p, const farray& a) { upload4(p.getNetForce(), a); })
        .def("getNetForce", [](ParticleData& p) { return download4(p.getNetForce()); })
        .def("setNetTorque", [](ParticleData& p, const farray& a) { upload4(p.getNetTorqueArray(), a); })
        .def("getNetTorque", [](ParticleData& p) { return download4(p.getNetTorqueArray()); })
        .def("setExternalEnergy", &ParticleData::setExternalEnergy)
        .def("setPressureFlag", &ParticleData::setPressureFlag)
        .def("getPressureFlag", &ParticleData::getPressureFlag)
        .def("getExternalVirial", &ParticleData::getExternalVirial)
        .def("setExternalVirial", &ParticleData::setExternalVirial)
        .def("positionsPointer", [](ParticleData& p) { return (size_t)p.getPositions().data(); });
    py::class_<SystemDefinition, std::shared_ptr<SystemDefinition>>(m, "SystemDefinition")
        .def(py::init<unsigned, const BoxDim&, const std::vector<std::string>&>())
        .def("getParticleData", &SystemDefinition::getParticleData)
        .def("getExecConf", &SystemDefinition::getExecConf);
    py::class_<ForceCompute, std::shared_ptr<ForceCompute>>(m, "ForceCompute")
        .def("compute", &ForceCompute::compute)
        .def("getForces", [](ForceCompute& f) { return download4(f.getForceArray()); })
        .def("getExternalVirial", &ForceCompute::getExternalVirial)
        .def("getProvidedLogQuantities", &ForceCompute::getProvidedLogQuantities)
        .def("getLogValue", &ForceCompute::getLogValue)
        .def_readwrite("enabled", &ForceCompute::enabled);
    py::class_<PrescribedForce, ForceCompute, std::shared_ptr<PrescribedForce>>(m, "PrescribedForce")
        .def(py::init<std::shared_ptr<SystemDefinition>>())
        .def("setArrays", [](PrescribedForce& f, const farray& force, const farray& torque, const farray& virial, Scalar ext) {
            auto v4 = [](const farray& a) { std::vector<Scalar4> o(a.shape(0)); memcpy(o.data(), a.data(), sizeof(Scalar4) * o.size()); return o; };
            std::vector<Scalar> v(virial.data(), virial.data() + virial.size());
            f.setArrays(v4(force), v4(torque), v, ext);
        })
        .def("getVirialPitch", &PrescribedForce::getVirialPitch)
        .def("getTorques", [](ForceCompute& f) { return download4(f.getTorqueArray()); })
        .def("getVirial", [](ForceCompute& f) {
            auto& a = f.getVirialArray();
            py::array_t<float> out((py::ssize_t)a.size());
            cuda_check(cudaDeviceSynchronize(), "sync");
            if (a.size()) a.download(out.mutable_data(), a.size());
            return out;
        });
    py::class_<System, std::shared_ptr<System>>(m, "System")
        .def(py::init<std::shared_ptr<SystemDefinition>>())
        .def("addCompute", &System::addCompute);

    // ---- CollectiveVariable (CollectiveVariable.cc:109-130) ----------------------------------------------------
    py::class_<CollectiveVariable, ForceCompute, std::shared_ptr<CollectiveVariable>> collective_variable(m, "CollectiveVariable");
    collective_variable.def(py::init<std::shared_ptr<SystemDefinition>, const std::string&>())
        .def("getCurrentValue", &CollectiveVariable::getCurrentValue)
        .def("setBiasFactor", &CollectiveVariable::setBiasFactor)
        .def("getBiasFactor", &CollectiveVariable::getBiasFactor)
        .def("setUmbrella", &CollectiveVariable::setUmbrella)
        .def("setKappa", &CollectiveVariable::setKappa)
        .def("setWidthFlat", &CollectiveVariable::setWidthFlat)
        .def("setMinimum", &CollectiveVariable::setMinimum)
        .def("setScale", &CollectiveVariable::setScale)
        .def("getName", &CollectiveVariable::getName)
        .def("getUmbrellaPotential", &CollectiveVariable::getUmbrellaPotential)
        .def("computeDerivatives", &CollectiveVariable::computeDerivatives)
        .def("canComputeDerivatives", &CollectiveVariable::canComputeDerivatives)
        .def("requiresNetForce", &CollectiveVariable::requiresNetForce);
    py::enum_<CollectiveVariable::umbrella_Enum>(collective_variable, "umbrella")
        .value("no_umbrella", CollectiveVariable::no_umbrella)
        .value("linear", CollectiveVariable::linear)
        .value("harmonic", CollectiveVariable::harmonic)
        .value("wall", CollectiveVariable::wall)
        .value("gaussian", CollectiveVariable::gaussian)
        .export_values();

    // ---- the CVs; the GPU classes ARE the implementation, the CPU class names alias them ------------------------
    py::class_<LamellarOrderParameterGPU, CollectiveVariable, std::shared_ptr<LamellarOrderParameterGPU>> lam(m, "LamellarOrderParameterGPU");
    lam.def(py::init<std::shared_ptr<SystemDefinition>, const std::vector<Scalar>&, const std::vector<int3_>&, const std::string&>())
        .def_readwrite("allreduce", &LamellarOrderParameterGPU::allreduce);
    m.attr("LamellarOrderParameter") = lam;
    py::class_<OrderParameterMeshGPU, CollectiveVariable, std::shared_ptr<OrderParameterMeshGPU>> mesh(m, "OrderParameterMeshGPU");
    mesh.def(py::init<std::shared_ptr<SystemDefinition>, unsigned int, unsigned int, unsigned int, std::vector<Scalar>, std::vector<int3_>>())
        .def("setTable", &OrderParameterMeshGPU::setTable)
        .def("setUseTable", &OrderParameterMeshGPU::setUseTable);
    m.attr("OrderParameterMesh") = mesh;
    py::class_<WellTemperedEnsemble, CollectiveVariable, std::shared_ptr<WellTemperedEnsemble>>(m, "WellTemperedEnsemble")
        .def(py::init<std::shared_ptr<SystemDefinition>, const std::string&>());
    py::class_<AspectRatio, CollectiveVariable, std::shared_ptr<AspectRatio>>(m, "AspectRatio")
        .def(py::init<std::shared_ptr<SystemDefinition>, const unsigned int, const unsigned int>());
    py::class_<Density, CollectiveVariable, std::shared_ptr<Density>>(m, "Density")
        .def(py::init<std::shared_ptr<SystemDefinition>, const std::string&>());

    py::class_<CollectiveWrapper, CollectiveVariable, std::shared_ptr<CollectiveWrapper>>(m, "CollectiveWrapper")
        .def(py::init<std::shared_ptr<SystemDefinition>, std::shared_ptr<ForceCompute>, const std::string&>())
        .def_readwrite("allreduce", &CollectiveWrapper::allreduce);

    py::class_<IndexGrid>(m, "IndexGrid")
        .def(py::init<const std::vector<unsigned int>&>())
        .def("getIndex", &IndexGrid::getIndex)
        .def("getCoordinates", [](IndexGrid& g, unsigned int idx) { std::vector<unsigned int> c(g.getDimension()); g.getCoordinates(idx, c); return c; })
        .def("getNumElements", &IndexGrid::getNumElements);

    // ---- IntegratorMetaDynamics (IntegratorMetaDynamics.cc:1315-1349) ---------------------------------------------
    py::class_<IntegratorMetaDynamics, std::shared_ptr<IntegratorMetaDynamics>> integrator_metad(m, "IntegratorMetaDynamics");
    integrator_metad
        .def(py::init<std::shared_ptr<SystemDefinition>, Scalar, Scalar, Scalar, Scalar, unsigned int, bool, const std::string&, bool,
                      IntegratorMetaDynamics::Enum>())
        .def("registerCollectiveVariable", &IntegratorMetaDynamics::registerCollectiveVariable)
        .def("removeAllVariables", &IntegratorMetaDynamics::removeAllVariables)
        .def("isInitialized", &IntegratorMetaDynamics::isInitialized)
        .def("setGrid", &IntegratorMetaDynamics::setGrid)
        .def("dumpGrid", &IntegratorMetaDynamics::dumpGrid)
        .def("restartFromGridFile", &IntegratorMetaDynamics::restartFromGridFile)
        .def("setAddHills", &IntegratorMetaDynamics::setAddHills)
        .def("setMode", &IntegratorMetaDynamics::setMode)
        .def("setStride", &IntegratorMetaDynamics::setStride)
        .def("setAdaptive", &IntegratorMetaDynamics::setAdaptive)
        .def("setSigmaG", &IntegratorMetaDynamics::setSigmaG)
        .def("resetHistogram", &IntegratorMetaDynamics::resetHistogram)
        .def("setMultipleWalkers", &IntegratorMetaDynamics::setMultipleWalkers)
        .def("setSystem", &IntegratorMetaDynamics::setSystem)
        .def("prepRun", &IntegratorMetaDynamics::prepRun)
        .def("update", &IntegratorMetaDynamics::update)
        .def("getProvidedLogQuantities", &IntegratorMetaDynamics::getProvidedLogQuantities)
        .def("getLogValue", &IntegratorMetaDynamics::getLogValue)
        .def("getGridArray", &IntegratorMetaDynamics::getGridArray)
        .def("getNumGaussians", &IntegratorMetaDynamics::getNumGaussians)
        .def("getSigmaInv", &IntegratorMetaDynamics::getSigmaInv)
        .def_readwrite("walker_allreduce", &IntegratorMetaDynamics::walker_allreduce)
        .def_readwrite("domain_allreduce", &IntegratorMetaDynamics::domain_allreduce);
    py::enum_<IntegratorMetaDynamics::Enum>(integrator_metad, "mode")
        .value("standard", IntegratorMetaDynamics::mode_standard)
        .value("well_tempered", IntegratorMetaDynamics::mode_well_tempered)
        .export_values();
}
