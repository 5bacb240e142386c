"""Collective variables -- the reference's `metadynamics.cv` Python API (cv.py), same class names, keyword
arguments, defaults and error behaviour, bound to the B200-native `_metadynamics` module.

    cv.lamellar(mode, lattice_vectors, name=None, sigma=1.0)                      reference cv.py:173-275
    cv.mesh(mode, nx, ny=None, nz=None, name=None, sigma=1.0, zero_modes=None)    reference cv.py:354-470
    cv.potential_energy(sigma=1.0)                                                reference cv.py:473-501
    cv.aspect_ratio(dir1, dir2, name="", sigma=1.0)                               reference cv.py:277-306
    cv.density(group=None, sigma=1.0)                                             reference cv.py:309-339
    base class: set_grid(cv_min, cv_max, num_points), set_params(sigma, kappa, cv0, umbrella, width_flat, scale,
    reweight)                                                                     reference cv.py:11-171
    cv.wrap(force, sigma=1.0)                                                     reference cv.py:504-541
cv.steinhardt needs HOOMD's neighbour list and is out of scope (OUT OF SCOPE).
"""
# This build always binds to the stand-in for the handful of HOOMD-blue 2.x objects the scripts touch (hoomd_shim.py);
# binding the classes to a real HOOMD installation is a build-time step described in INTEGRATION.md.
from . import hoomd_shim as hoomd
from . import _metadynamics

_force_base = hoomd._force


class _collective_variable(_force_base):
    """Base class for collective variables (reference cv.py:11-171)."""

    def __init__(self, sigma, name=None):
        _force_base.__init__(self, name)
        self.sigma = sigma
        self.cv_min = 0.0
        self.cv_max = 0.0
        self.num_points = 0
        self.grid_set = False
        self.ftm_min = 0.0
        self.ftm_max = 0.0
        self.ftm_parameters_set = False
        self.umbrella = False
        self.reweight = False

    def set_grid(self, cv_min, cv_max, num_points):
        hoomd.util.print_status_line()
        self.cv_min = cv_min
        self.cv_max = cv_max
        self.num_points = int(num_points)
        self.grid_set = True

    def enable_histograms(self, ftm_min, ftm_max):
        hoomd.util.print_status_line()
        self.ftm_min = ftm_min
        self.ftm_max = ftm_max
        self.ftm_parameters_set = True

    def set_params(self, sigma=None, kappa=None, cv0=None, umbrella=None, width_flat=None, scale=None, reweight=None):
        hoomd.util.print_status_line()
        if sigma is not None:
            self.sigma = sigma
        if umbrella is not None:
            modes = {"no_umbrella": (self.cpp_force.umbrella.no_umbrella, False), "linear": (self.cpp_force.umbrella.linear, True),
                     "harmonic": (self.cpp_force.umbrella.harmonic, True), "wall": (self.cpp_force.umbrella.wall, True),
                     "gaussian": (self.cpp_force.umbrella.gaussian, True)}
            if umbrella not in modes:
                hoomd.context.msg.error("cv: Invalid umbrella mode specified.")
                raise RuntimeError("Error setting parameters of collective variable.")
            cpp_umbrella, on = modes[umbrella]
            self.reweight = on
            self.umbrella = on
            self.cpp_force.setUmbrella(cpp_umbrella)
        if kappa is not None:
            self.cpp_force.setKappa(kappa)
        if width_flat is not None:
            self.cpp_force.setWidthFlat(width_flat)
        if cv0 is not None:
            self.cpp_force.setMinimum(cv0)
        if scale is not None:
            self.cpp_force.setScale(scale)
        if reweight is not None:
            self.reweight = reweight

    def update_coeffs(self):
        pass


def _per_type_modes(mode, who):
    if type(mode) != type(dict()):
        hoomd.context.msg.error("cv.%s: Mode amplitudes specified incorrectly.\n" % who)
        raise RuntimeError('Error creating collective variable.')
    pdata = hoomd.context.current.system_definition.getParticleData()
    cpp_mode = hoomd.std_vector_scalar()
    for i in range(0, pdata.getNTypes()):
        t = pdata.getNameByType(i)
        if t not in mode.keys():
            hoomd.context.msg.error("cv.%s: Missing mode amplitude for particle type %s.\n" % (who, t))
            raise RuntimeError('Error creating collective variable.')
        cpp_mode.append(mode[t])
    return cpp_mode


def _int3_list(vectors, who):
    out = _metadynamics.std_vector_int3()
    for l in vectors:
        if len(l) != 3:
            hoomd.context.msg.error("cv.%s: List of input lattice vectors not a list of triples.\n" % who)
            raise RuntimeError('Error creating collective variable.')
        out.append(hoomd.make_int3(l[0], l[1], l[2]))
    return out


class lamellar(_collective_variable):
    """Lamellar order parameter s = (1/N) sum_k Re sum_j a(type_j) exp(i q_k.r_j) (reference cv.py:173-275)."""

    def __init__(self, mode, lattice_vectors, name=None, sigma=1.0):
        hoomd.util.print_status_line()
        if name is not None:
            name = "_" + name
            suffix = name
        else:
            suffix = ""
        _collective_variable.__init__(self, sigma, name)
        if len(lattice_vectors) == 0:
            hoomd.context.msg.error("cv.lamellar: List of supplied latice vectors is empty.\n")
            raise RuntimeError('Error creating collective variable.')
        cpp_mode = _per_type_modes(mode, "lamellar")
        cpp_lattice_vectors = _int3_list(lattice_vectors, "lamellar")
        # exec_conf.isCUDAEnabled() is always true here: the GPU class is the implementation (no CPU fallback)
        self.cpp_force = _metadynamics.LamellarOrderParameterGPU(
            hoomd.context.current.system_definition, cpp_mode, cpp_lattice_vectors, suffix)
        hoomd.context.current.system.addCompute(self.cpp_force, self.force_name)


class aspect_ratio(_collective_variable):
    """Aspect ratio L_dir1/L_dir2 of the box (reference cv.py:277-306)."""

    def __init__(self, dir1, dir2, name="", sigma=1.0):
        hoomd.util.print_status_line()
        _collective_variable.__init__(self, sigma, name)
        self.cpp_force = _metadynamics.AspectRatio(hoomd.context.current.system_definition, int(dir1), int(dir2))
        hoomd.context.current.system.addCompute(self.cpp_force, self.force_name)


class density(_collective_variable):
    """Number density N/V of all particles (reference cv.py:309-339; particle groups other than `all` need HOOMD)."""

    def __init__(self, group=None, sigma=1.0):
        hoomd.util.print_status_line()
        suffix = "" if group is None else str(getattr(group, "name", group))
        _collective_variable.__init__(self, sigma, "cv_density" + ("_" + suffix if suffix else ""))
        self.cpp_force = _metadynamics.Density(hoomd.context.current.system_definition, suffix)
        hoomd.context.current.system.addCompute(self.cpp_force, self.force_name)


class mesh(_collective_variable):
    """Particle-mesh structure-factor order parameter (reference cv.py:354-470)."""

    def __init__(self, mode, nx, ny=None, nz=None, name=None, sigma=1.0, zero_modes=None):
        hoomd.util.print_status_line()
        if name is not None:
            name = "_" + name
        if ny is None:
            ny = nx
        if nz is None:
            nz = nx
        _collective_variable.__init__(self, sigma, name)
        cpp_mode = _per_type_modes(mode, "mesh")
        cpp_zero_modes = _int3_list(zero_modes if zero_modes is not None else [], "mesh")
        self.cpp_force = _metadynamics.OrderParameterMeshGPU(
            hoomd.context.current.system_definition, int(nx), int(ny), int(nz), cpp_mode, cpp_zero_modes)
        hoomd.context.current.system.addCompute(self.cpp_force, self.force_name)

    def set_params(self, use_table=None, **args):
        hoomd.util.print_status_line()
        if use_table is not None:
            self.cpp_force.setUseTable(use_table)
        hoomd.util.quiet_status()
        _collective_variable.set_params(self, **args)
        hoomd.util.unquiet_status()

    def set_kernel(self, func, kmin, kmax, width, coeff=dict()):
        Ktable = hoomd.std_vector_scalar()
        dKtable = hoomd.std_vector_scalar()
        dk = (kmax - kmin) / float(width - 1)
        for i in range(0, width):
            k = kmin + dk * i
            (K, dK) = func(k, kmin, kmax, **coeff)
            Ktable.append(K)
            dKtable.append(dK)
        self.cpp_force.setTable(Ktable, dKtable, kmin, kmax)


class potential_energy(_collective_variable):
    """Potential energy as a collective variable: well-tempered ensemble (reference cv.py:473-501)."""

    def __init__(self, sigma=1.0):
        hoomd.util.print_status_line()
        name = 'cv_potential_energy'
        _collective_variable.__init__(self, sigma, name)
        self.enabled = False                      # disable as regular ForceCompute (cv.py:490)
        self.cpp_force = _metadynamics.WellTemperedEnsemble(hoomd.context.current.system_definition, name)
        self.cpp_force.enabled = False
        hoomd.context.current.system.addCompute(self.cpp_force, name)


class wrap(_collective_variable):
    """Force wrapper: use the potential energy of an arbitrary force as collective variable (reference cv.py:504-541).
    `force` is a force object of the MD engine (`md.force._force`; with the stand-in: hoomd_shim.prescribed_force).  The
    reference's disable() / enable() recurse into themselves and name an undefined variable (cv.py:531-537); here they do
    what they were written for: switch this CV and the wrapped force together."""

    def __init__(self, force, sigma=1.0):
        hoomd.util.print_status_line()
        if not isinstance(force, _force_base):
            hoomd.context.msg.error("cv.wrap needs a md._force instance as argument.")
            raise RuntimeError("Error creating cv.wrap")
        name = 'cv_' + force.name
        _collective_variable.__init__(self, sigma, name)
        self.force = force
        self.cpp_force = _metadynamics.CollectiveWrapper(hoomd.context.current.system_definition, force.cpp_force, name)
        if force.enabled or force.log:
            hoomd.context.current.system.addCompute(self.cpp_force, name)
        self.log = force.log

    def disable(self, log=False):
        _collective_variable.disable(self, log)
        self.force.disable(log)

    def enable(self):
        _collective_variable.enable(self)
        self.force.enable()
