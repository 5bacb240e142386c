"""Minimal stand-in for the pieces of the `hoomd` 2.x Python package that the plugin's Python layer touches
(reference cv.py:2-8, integrate.py:1-30): `context.current.{system_definition,system,forces}`,
`context.exec_conf`, `context.msg`, `util.print_status_line/quiet_status`, the `_force` / `_integrator` base
classes, `make_int3`, `std_vector_scalar`, and a `run()` loop that drives the integrator the way HOOMD's
System::run does (prepRun once, then update() per step).  With a real HOOMD installed the same cv.py /
integrate.py logic binds to hoomd itself."""
import numpy as np

from . import _metadynamics as _m


class _Msg:
    def __init__(self):
        self.warnings, self.errors = [], []

    def error(self, s):
        self.errors.append(s)

    def warning(self, s):
        self.warnings.append(s)

    def notice(self, level, s):
        pass


class _Current:
    def __init__(self):
        self.system_definition = None
        self.system = None
        self.forces = []
        self.integrator = None
        self.timestep = 0


class _ExecConf:
    def isCUDAEnabled(self):
        return True


class _Context:
    def __init__(self):
        self.current = _Current()
        self.msg = _Msg()
        self.exec_conf = _ExecConf()

    def initialize(self):
        self.current = _Current()
        self.msg = _Msg()
        return self


context = _Context()


class util:
    @staticmethod
    def print_status_line():
        pass

    @staticmethod
    def quiet_status():
        pass

    @staticmethod
    def unquiet_status():
        pass


def make_int3(x, y, z):
    return _m.make_int3(int(x), int(y), int(z))


std_vector_scalar = _m.std_vector_scalar


class init:
    @staticmethod
    def from_arrays(positions, types, type_names, L, tilt=(0.0, 0.0, 0.0)):
        """Create the system from numpy arrays: positions (N,3), integer type ids (N,), type names, box lengths."""
        pos = np.asarray(positions, dtype=np.float32)
        n = pos.shape[0]
        try:
            lx, ly, lz = (float(v) for v in L)
        except TypeError:
            lx = ly = lz = float(L)
        box = _m.BoxDim(lx, ly, lz, *[float(t) for t in tilt])
        sysdef = _m.SystemDefinition(n, box, list(type_names))
        pt = np.empty((n, 4), dtype=np.float32)
        pt[:, :3] = pos
        pt[:, 3] = np.asarray(types, dtype=np.int32).view(np.float32)
        sysdef.getParticleData().setPositions(pt)
        context.current.system_definition = sysdef
        context.current.system = _m.System(sysdef)
        return sysdef


class _force:
    """hoomd.md.force._force: registers itself in context.current.forces under a unique name."""
    _count = 0

    def __init__(self, name=None):
        if context.current.system is None:
            context.msg.error("Cannot create force before initialization\n")
            raise RuntimeError("Error creating force")
        suffix = "" if name is None else "_" + name
        self.force_name = "force%d%s" % (_force._count, suffix)
        self.name = self.force_name if name is None else name
        _force._count += 1
        self.enabled = True
        self.log = True
        self.cpp_force = None
        context.current.forces.append(self)

    def disable(self, log=False):
        self.enabled = False
        self.cpp_force.enabled = False

    def enable(self):
        self.enabled = True
        self.cpp_force.enabled = True

    def get_forces(self):
        return self.cpp_force.getForces()


class prescribed_force(_force):
    """Stand-in for "any force of the MD engine" (pair, bond, ...): per-particle force / energy (N,4), torque (N,4), virial
    (6,N) and an external energy given as arrays; what cv.wrap wraps when HOOMD-blue itself is absent."""

    def __init__(self, force4, torque4=None, virial6N=None, external_energy=0.0, name=None):
        import numpy as np
        from . import _metadynamics
        _force.__init__(self, name)
        sd = context.current.system_definition
        f = np.ascontiguousarray(force4, dtype=np.float32)
        n = f.shape[0]
        self.cpp_force = _metadynamics.PrescribedForce(sd)
        pitch = self.cpp_force.getVirialPitch()
        t = np.zeros((n, 4), np.float32) if torque4 is None else np.ascontiguousarray(torque4, dtype=np.float32)
        v = np.zeros((6, pitch), np.float32)
        if virial6N is not None:
            v[:, :n] = np.asarray(virial6N, dtype=np.float32)
        self.cpp_force.setArrays(f, t, v.reshape(-1), float(external_energy))
        context.current.system.addCompute(self.cpp_force, self.force_name)


class _integrator:
    """hoomd.md.integrate._integrator."""

    def __init__(self):
        self.cpp_integrator = None
        self.supports_methods = False
        context.current.integrator = self

    def update_forces(self):
        for f in context.current.forces:
            if f.cpp_force is not None:
                f.cpp_force.enabled = f.enabled


def run(tsteps):
    """HOOMD's run(): update_forces(), prepRun(timestep) once per run call, then update(timestep) per step."""
    cur = context.current
    integ = cur.integrator
    if integ is None:
        raise RuntimeError("No integrator set")
    integ.update_forces()
    integ.cpp_integrator.setSystem(cur.system)
    integ.cpp_integrator.prepRun(cur.timestep)
    for _ in range(int(tsteps)):
        integ.cpp_integrator.update(cur.timestep)
        cur.timestep += 1
