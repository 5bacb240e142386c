"""Metadynamics integration mode -- the reference's `metadynamics.integrate` Python API (integrate.py:204-357),
same signatures and defaults, bound to the B200-native `_metadynamics.IntegratorMetaDynamics`."""
# This build always binds to the stand-in for the handful of HOOMD-blue 2.x objects the scripts touch (hoomd_shim.py);
# binding the classes to a real HOOMD installation is a build-time step described in INTEGRATION.md.
from . import hoomd_shim as hoomd
from . import _metadynamics
from . import cv


class mode_metadynamics(hoomd._integrator):
    """integrate.mode_metadynamics(dt, stride, mode="standard", W=1.0, deltaT=1.0, T=1.0, filename="",
    overwrite=False, add_hills=True) -- reference integrate.py:204-226."""

    def __init__(self, dt, stride, mode="standard", W=1.0, deltaT=1.0, T=1.0, filename="", overwrite=False, add_hills=True):
        hoomd.util.print_status_line()
        hoomd._integrator.__init__(self)
        if (mode == "standard"):
            cpp_mode = _metadynamics.IntegratorMetaDynamics.mode.standard
        elif (mode == "well_tempered"):
            cpp_mode = _metadynamics.IntegratorMetaDynamics.mode.well_tempered
        else:
            hoomd.context.msg.error("integrate.mode_metadynamics: Unsupported metadynamics mode.\n")
            raise RuntimeError('Error setting up Metadynamics.')
        self.cpp_integrator = _metadynamics.IntegratorMetaDynamics(
            hoomd.context.current.system_definition, dt, W, deltaT, T, int(stride), add_hills, filename, overwrite, cpp_mode)
        self.supports_methods = True
        self.cv_names = []

    def update_forces(self):
        """Registers the collective variables with the C++ integration class (reference integrate.py:228-269)."""
        if self.cpp_integrator.isInitialized():
            notfound = False
            num_cv = 0
            for f in hoomd.context.current.forces:
                if isinstance(f, cv._collective_variable) and f.grid_set:
                    if num_cv >= len(self.cv_names) or f.name != self.cv_names[num_cv]:
                        notfound = True
                    num_cv += 1
            if (len(self.cv_names) != num_cv) or notfound:
                hoomd.context.msg.error(
                    "integrate.mode_metadynamics: Set of collective variables has changed since last run. This is unsupported.\n")
                raise RuntimeError('Error setting up Metadynamics.')
        self.cv_names = []
        self.cpp_integrator.removeAllVariables()
        for f in hoomd.context.current.forces:
            if isinstance(f, cv._collective_variable):
                if f.grid_set is True:
                    self.cpp_integrator.registerCollectiveVariable(f.cpp_force, f.sigma, f.cv_min, f.cv_max, f.num_points)
                    self.cv_names.append(f.name)
                else:
                    if not f.umbrella:
                        hoomd.context.msg.warning("integrate.mode_metadynamics: Grid parameters not set. Ignoring CV " + f.name)
        if len(self.cv_names) == 0:
            hoomd.context.msg.warning(
                "integrate.mode_metadynamics: No collective variables defined. Continuing with simulation anyway.\n")
        if not self.cpp_integrator.isInitialized():
            self.cpp_integrator.setGrid(True)
        hoomd._integrator.update_forces(self)

    def dump_grid(self, filename1, filename2="", period=0):
        hoomd.util.print_status_line()
        self.cpp_integrator.dumpGrid(filename1, filename2, int(period))

    def restart_from_grid(self, filename):
        hoomd.util.print_status_line()
        self.cpp_integrator.restartFromGridFile(filename)

    def reset_histogram(self):
        hoomd.util.print_status_line()
        self.cpp_integrator.resetHistogram()

    def set_params(self, add_hills=None, mode=None, stride=None, adaptive=None, sigma_g=None, multiple_walkers=None):
        hoomd.util.print_status_line()
        if add_hills is not None:
            self.cpp_integrator.setAddHills(add_hills)
        if mode is not None:
            if (mode == "standard"):
                cpp_mode = _metadynamics.IntegratorMetaDynamics.mode.standard
            elif (mode == "well_tempered"):
                cpp_mode = _metadynamics.IntegratorMetaDynamics.mode.well_tempered
            else:
                hoomd.context.msg.error("integrate.mode_metadynamics: Unsupported metadynamics mode.\n")
                raise RuntimeError('Error setting up Metadynamics.')
            self.cpp_integrator.setMode(cpp_mode)
        if stride is not None:
            self.cpp_integrator.setStride(int(stride))
        if adaptive is not None:
            self.cpp_integrator.setAdaptive(adaptive)
        if sigma_g is not None:
            self.cpp_integrator.setSigmaG(sigma_g)
        if multiple_walkers is not None:
            self.cpp_integrator.setMultipleWalkers(multiple_walkers)


class mode_standard(hoomd._integrator):
    """Stand-in for hoomd.md.integrate.mode_standard (reference test/test_mesh.py:16): no bias grid, every enabled
    force compute is evaluated once per step through CollectiveVariable::computeForces (umbrella potentials)."""

    class _Cpp:
        def __init__(self):
            self.system = None
            self.t_last = None

        def setSystem(self, system):
            self.system = system

        def prepRun(self, timestep):
            pass

        def update(self, timestep):
            for f in hoomd.context.current.forces:
                if f.enabled and f.cpp_force is not None:
                    f.cpp_force.compute(timestep + 1)

    def __init__(self, dt):
        hoomd._integrator.__init__(self)
        self.dt = dt
        self.cpp_integrator = mode_standard._Cpp()
