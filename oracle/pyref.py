"""ctypes access to oracle/_ref/libref_{f32,f64}.so: the REFERENCE's own CPU classes (CollectiveVariable, LamellarOrderParameter,
OrderParameterMesh, AspectRatio, IndexGrid), compiled from the reference's sources where they lie against the HOOMD stand-in in
oracle/ref_shim/ (see oracle/Makefile target `ref`, oracle/ref_capi.cc).  TEST INFRASTRUCTURE ONLY: used by
tests/test_reference_build.py and tests/golden/make_ref_golden.py to pin the oracle's restatement."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = "/root/reference/metadynamics"
_libs = {}
_dp, _fp, _ip, _up = C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_uint)


def available():
    return os.path.isdir(REFERENCE) or all(os.path.exists(os.path.join(_HERE, "_ref", "libref_%s.so" % p)) for p in ("f32", "f64"))


def build():
    if os.path.isdir(REFERENCE):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


def lib(prec):
    if prec not in _libs:
        build()
        _libs[prec] = C.CDLL(os.path.join(_HERE, "_ref", "libref_%s.so" % prec))
        assert _libs[prec].ref_scalar_bytes() == (4 if prec == "f32" else 8)
    return _libs[prec]


def _d(a):
    return a.ctypes.data_as(_dp)


def _box(L, tilt):
    return (np.ascontiguousarray(np.broadcast_to(np.asarray(L, np.float64), (3,))), np.ascontiguousarray(np.asarray(tilt, np.float64)))


def mesh(dims, mode, L, postype, bias, prec="f64", tilt=(0, 0, 0)):
    nx, ny, nz = dims
    M, N = nx * ny * nz, postype.shape[0]
    mode = np.ascontiguousarray(mode, np.float64)
    Lb, tb = _box(L, tilt)
    pt = np.ascontiguousarray(postype, np.float32)
    cv, msq = C.c_double(), C.c_double()
    force, rho, inv, interp = np.empty((N, 4)), np.empty(M), np.empty(M), np.empty(M)
    rc = lib(prec).ref_mesh(nx, ny, nz, _d(mode), len(mode), _d(Lb), _d(tb), pt.ctypes.data_as(_fp), N, C.c_double(bias), C.byref(cv), C.byref(msq),
                            _d(force), _d(rho), _d(inv), _d(interp))
    assert rc == 0
    return dict(cv=cv.value, mode_sq=msq.value, force=force, rho=rho.reshape(nz, ny, nx), inv=inv.reshape(nz, ny, nx),
                interp=interp.reshape(nz, ny, nx))


def mesh_virial(dims, mode, L, postype, K, dK, kmin, kmax, bias, prec="f64", tilt=(0, 0, 0)):
    """OrderParameterMesh::computeVirial of the reference with a kernel table: external virial (6)."""
    nx, ny, nz = dims
    mode = np.ascontiguousarray(mode, np.float64)
    Lb, tb = _box(L, tilt)
    pt = np.ascontiguousarray(postype, np.float32)
    K, dK = np.ascontiguousarray(K, np.float64), np.ascontiguousarray(dK, np.float64)
    out = np.empty(6)
    rc = lib(prec).ref_mesh_virial(nx, ny, nz, _d(mode), len(mode), _d(Lb), _d(tb), pt.ctypes.data_as(_fp), pt.shape[0], _d(K), _d(dK), len(K),
                                   C.c_double(kmin), C.c_double(kmax), C.c_double(bias), _d(out))
    assert rc == 0
    return out


def mesh_qmax(dims, mode, L, postype, prec="f64", tilt=(0, 0, 0)):
    """OrderParameterMesh::computeQmax of the reference: array {q_max.x, q_max.y, q_max.z, sq_max}."""
    nx, ny, nz = dims
    mode = np.ascontiguousarray(mode, np.float64)
    Lb, tb = _box(L, tilt)
    pt = np.ascontiguousarray(postype, np.float32)
    out = np.empty(4)
    rc = lib(prec).ref_mesh_qmax(nx, ny, nz, _d(mode), len(mode), _d(Lb), _d(tb), pt.ctypes.data_as(_fp), pt.shape[0], _d(out))
    assert rc == 0
    return out


def lamellar(mode, lattice_vectors, L, postype, bias, prec="f64", tilt=(0, 0, 0)):
    N = postype.shape[0]
    mode = np.ascontiguousarray(mode, np.float64)
    lv = np.ascontiguousarray(lattice_vectors, np.int32).reshape(-1, 3)
    Lb, tb = _box(L, tilt)
    pt = np.ascontiguousarray(postype, np.float32)
    cv = C.c_double()
    modes, force = np.empty(2 * lv.shape[0]), np.empty((N, 4))
    rc = lib(prec).ref_lamellar(_d(mode), len(mode), lv.ctypes.data_as(_ip), lv.shape[0], _d(Lb), _d(tb), pt.ctypes.data_as(_fp), N,
                                C.c_double(bias), C.byref(cv), _d(modes), _d(force))
    assert rc == 0
    return dict(cv=cv.value, modes=modes.reshape(-1, 2), force=force)


def umbrella(kind, mode, lattice_vectors, L, postype, bias_in=0.0, cv0=0.0, kappa=1.0, width_flat=0.0, scale=1.0, prec="f64"):
    N = postype.shape[0]
    mode = np.ascontiguousarray(mode, np.float64)
    lv = np.ascontiguousarray(lattice_vectors, np.int32).reshape(-1, 3)
    Lb, _ = _box(L, (0, 0, 0))
    pt = np.ascontiguousarray(postype, np.float32)
    cv, en = C.c_double(), C.c_double()
    force = np.empty((N, 4))
    kinds = dict(no_umbrella=0, linear=1, harmonic=2, wall=3, gaussian=4)
    rc = lib(prec).ref_umbrella(kinds[kind], C.c_double(cv0), C.c_double(kappa), C.c_double(width_flat), C.c_double(scale), _d(mode), len(mode),
                                lv.ctypes.data_as(_ip), lv.shape[0], _d(Lb), pt.ctypes.data_as(_fp), N, C.c_double(bias_in), C.byref(cv), C.byref(en),
                                _d(force))
    assert rc == 0
    return dict(cv=cv.value, energy=en.value, force=force)


def aspect(dir1, dir2, L, bias, prec="f64", tilt=(0, 0, 0)):
    Lb, tb = _box(L, tilt)
    cv = C.c_double()
    vir = np.empty(6)
    assert lib(prec).ref_aspect(dir1, dir2, _d(Lb), _d(tb), C.c_double(bias), C.byref(cv), _d(vir)) == 0
    return cv.value, vir


def indexgrid_index(lengths, coords):
    l, c = np.ascontiguousarray(lengths, np.uint32), np.ascontiguousarray(coords, np.uint32)
    f = lib("f64").ref_indexgrid_index
    f.restype = C.c_uint
    return int(f(l.ctypes.data_as(_up), len(l), c.ctypes.data_as(_up)))


def indexgrid_coords(lengths, idx):
    l = np.ascontiguousarray(lengths, np.uint32)
    out = np.empty(len(l), np.uint32)
    lib("f64").ref_indexgrid_coords(l.ctypes.data_as(_up), len(l), C.c_uint(idx), out.ctypes.data_as(_up))
    return out


def indexgrid_num(lengths):
    l = np.ascontiguousarray(lengths, np.uint32)
    f = lib("f64").ref_indexgrid_num
    f.restype = C.c_uint
    return int(f(l.ctypes.data_as(_up), len(l)))


def grid_sequence(cv_min, cv_max, num_points, sigma, cv_values, timesteps, W=1.0, T_shift=1.0, T=1.0, stride=1, add_bias=True,
                  well_tempered=False, prec="f64"):
    """IntegratorMetaDynamics: prepRun + updateBiasPotential per step with prescribed CV values.  Returns the bias factors
    after every step and the final grid state."""
    a, b, s_ = (np.ascontiguousarray(v, np.float64) for v in (cv_min, cv_max, sigma))
    n = np.ascontiguousarray(num_points, np.uint32)
    d = len(n)
    G = int(np.prod(n))
    vals = np.ascontiguousarray(cv_values, np.float64).reshape(-1, d)
    ts = np.ascontiguousarray(timesteps, np.uint32)
    bias = np.empty_like(vals)
    out = {k: np.empty(G) for k in ("grid", "reweighted", "weight", "sigma_grid")}
    outu = {k: np.empty(G, np.uint32) for k in ("hist", "hist_gauss", "hist_delta")}
    sc = np.empty(3)
    rc = lib(prec).ref_grid_sequence(d, _d(a), _d(b), n.ctypes.data_as(_up), _d(s_), C.c_double(W), C.c_double(T_shift), C.c_double(T),
                                     C.c_uint(stride), int(add_bias), int(well_tempered), _d(vals), ts.ctypes.data_as(_up), len(ts), _d(bias),
                                     _d(out["grid"]), _d(out["reweighted"]), _d(out["weight"]), _d(out["sigma_grid"]),
                                     outu["hist"].ctypes.data_as(_up), outu["hist_gauss"].ctypes.data_as(_up), outu["hist_delta"].ctypes.data_as(_up), _d(sc))
    assert rc == 0
    out.update(outu)
    out.update(bias=bias, bias_potential=sc[0], reweight=sc[1], num_gaussians=int(sc[2]))
    return out


def grid_adaptive(cv_min, cv_max, num_points, sigma, grads, can, sigma_g, cv_values, timesteps, W=1.0, T_shift=1.0, T=1.0, stride=1,
                  well_tempered=False, prec="f64"):
    """IntegratorMetaDynamics with adaptive Gaussians: prescribed CV values and gradients (grads: (ncv, N, 4) float32)."""
    a, b, s_ = (np.ascontiguousarray(v, np.float64) for v in (cv_min, cv_max, sigma))
    n = np.ascontiguousarray(num_points, np.uint32)
    d = len(n)
    G = int(np.prod(n))
    g = np.ascontiguousarray(grads, np.float32)
    N = g.shape[1]
    cani = np.ascontiguousarray(can, np.int32)
    vals = np.ascontiguousarray(cv_values, np.float64).reshape(-1, d)
    ts = np.ascontiguousarray(timesteps, np.uint32)
    bias = np.empty_like(vals)
    sinv = np.empty((len(ts), d, d))
    grid, sgrid = np.empty(G), np.empty(G)
    rc = lib(prec).ref_grid_adaptive(d, _d(a), _d(b), n.ctypes.data_as(_up), _d(s_), C.c_double(W), C.c_double(T_shift), C.c_double(T),
                                     C.c_uint(stride), int(well_tempered), C.c_double(sigma_g), g.ctypes.data_as(C.POINTER(C.c_float)),
                                     cani.ctypes.data_as(C.POINTER(C.c_int)), C.c_uint(N), _d(vals), ts.ctypes.data_as(_up), len(ts),
                                     _d(bias), _d(sinv), _d(grid), _d(sgrid))
    assert rc == 0
    return dict(bias=bias, sigma_inv=sinv, grid=grid, sigma_grid=sgrid)


def test2d_files(directory, restart=False, prec="f64"):
    """The reference's test/test_2d.py scenario through the reference's own IntegratorMetaDynamics / Density / AspectRatio,
    including the files they write into `directory` (grid dumps, hills log); returns num_gaussians."""
    rc = lib(prec).ref_test2d_files(os.fsencode(directory), 1 if restart else 0)
    assert rc >= 0
    return rc


def wte(net_force, net_torque, net_virial6N, external_energy, external_virial6, bias, prec="f64"):
    """The reference's WellTemperedEnsemble (CPU branch): CV = sum net_force.w + external energy, then computeBiasForces.
    net_force / net_torque: (N, 4); net_virial6N: (6, N).  Returns dict(pe, force, torque, virial, external_virial)."""
    f = np.array(net_force, dtype=np.float64, copy=True, order="C")
    t = np.array(net_torque, dtype=np.float64, copy=True, order="C")
    v = np.array(net_virial6N, dtype=np.float64, copy=True, order="C")
    e = np.array(external_virial6, dtype=np.float64, copy=True)
    pe = C.c_double()
    rc = lib(prec).ref_wte(f.shape[0], _d(f), _d(t), _d(v), C.c_double(external_energy), _d(e), C.c_double(bias), C.byref(pe))
    assert rc == 0
    return dict(pe=pe.value, force=f, torque=t, virial=v, external_virial=e)


class StepPlan:
    """A CV object of the reference (OrderParameterMesh or LamellarOrderParameter) on a resident particle set; step(bias) runs
    getCurrentValue + setBiasFactor + computeBiasForces for a new timestep -- the reference's own CPU implementation of the
    hot path, what bench.py's `--impl reference` arm and `cpu_baseline` time when oracle/_ref is built."""

    def __init__(self, kind, postype, L, mode, prec="f32", dims=None, lattice_vectors=None, tilt=(0, 0, 0)):
        self.l = lib(prec)
        self.l.ref_step_create_mesh.restype = C.c_void_p
        self.l.ref_step_create_lamellar.restype = C.c_void_p
        self.l.ref_step.argtypes = [C.c_void_p, C.c_double, _dp, _dp]
        self.l.ref_step_force.argtypes = [C.c_void_p, C.c_uint, _dp]
        self.l.ref_step_destroy.argtypes = [C.c_void_p]
        mode = np.ascontiguousarray(mode, np.float64)
        Lb, tb = _box(L, tilt)
        pt = np.ascontiguousarray(postype, np.float32)
        if kind == "mesh":
            nx, ny, nz = dims
            self.h = self.l.ref_step_create_mesh(nx, ny, nz, _d(mode), len(mode), _d(Lb), _d(tb), pt.ctypes.data_as(_fp), pt.shape[0])
        else:
            lv = np.ascontiguousarray(lattice_vectors, np.int32).reshape(-1, 3)
            self.h = self.l.ref_step_create_lamellar(_d(mode), len(mode), lv.ctypes.data_as(_ip), lv.shape[0], _d(Lb), _d(tb),
                                                     pt.ctypes.data_as(_fp), pt.shape[0])
        if not self.h:
            raise RuntimeError("the reference's CV object could not be created")

    def step(self, bias):
        cv, sec = C.c_double(), np.zeros(2)
        if self.l.ref_step(self.h, float(bias), C.byref(cv), _d(sec)) != 0:
            raise RuntimeError("ref_step failed")
        return cv.value, float(sec[0]), float(sec[1])

    def force(self, i):
        out = np.zeros(4)
        self.l.ref_step_force(self.h, int(i), _d(out))
        return out

    def __del__(self):
        if getattr(self, "h", None):
            self.l.ref_step_destroy(self.h)
            self.h = None
