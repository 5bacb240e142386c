"""ctypes front-end of the CPU oracle (oracle/_build/liboracle.so).

TEST INFRASTRUCTURE ONLY -- pinned against the reference's own sources for the CV classes and the integrator
(see the header of oracle/metad_oracle.hpp and tests/test_reference_build.py).  Only tests/,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may import
this module; the product package (metadynamics_plugin_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")


def build(force=False):
    """Compile the oracle with the committed Makefile (g++, no external deps)."""
    src_m = max(os.path.getmtime(os.path.join(_HERE, f)) for f in ("capi.cc", "metad_oracle.hpp", "Makefile"))
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < src_m:
        subprocess.check_call(["make", "-C", _HERE, "clean", "all"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None

_dp = C.POINTER(C.c_double)
_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int)
_up = C.POINTER(C.c_uint)


def _d(a):
    return a.ctypes.data_as(_dp)


def _f(a):
    return a.ctypes.data_as(_fp)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        for sfx in ("f32", "f64"):
            g = lambda n: getattr(_lib, n + "_" + sfx)
            g("orc_mesh_create").restype = C.c_void_p
            g("orc_mesh_create").argtypes = [C.c_uint, C.c_uint, C.c_uint, _dp, C.c_int, _dp, C.c_uint]
            g("orc_mesh_destroy").argtypes = [C.c_void_p]
            g("orc_mesh_assign").argtypes = [C.c_void_p, _fp, C.c_uint]
            g("orc_mesh_update").argtypes = [C.c_void_p]
            g("orc_mesh_cv").restype = C.c_double
            g("orc_mesh_cv").argtypes = [C.c_void_p]
            g("orc_mesh_current_value").restype = C.c_double
            g("orc_mesh_current_value").argtypes = [C.c_void_p, _fp, C.c_uint]
            g("orc_mesh_forces").argtypes = [C.c_void_p, _fp, C.c_uint, C.c_double, _dp]
            g("orc_mesh_get").argtypes = [C.c_void_p, C.c_int, _dp]
            g("orc_mesh_cells").argtypes = [C.c_void_p, _ip, C.c_uint]
            g("orc_mesh_mode_sq").restype = C.c_double
            g("orc_mesh_mode_sq").argtypes = [C.c_void_p]
            g("orc_mesh_qmax").argtypes = [C.c_void_p, _dp]
            g("orc_mesh_virial").argtypes = [C.c_void_p, _dp, C.c_uint, C.c_double, C.c_double, C.c_int, C.c_double, _dp]
            g("orc_mesh_set_literal_copysignf").argtypes = [C.c_void_p, C.c_int]
            g("orc_mesh_set_literal_tilt_offset").argtypes = [C.c_void_p, C.c_int]
            g("orc_lamellar_cv").restype = C.c_double
            g("orc_lamellar_cv").argtypes = [_fp, C.c_uint, C.c_uint, _dp, C.c_int, _ip, C.c_int, _dp, _dp]
            g("orc_lamellar_forces").argtypes = [_fp, C.c_uint, C.c_uint, _dp, C.c_int, _ip, C.c_int, _dp, C.c_double, _dp]
            g("orc_grid_create").restype = C.c_void_p
            g("orc_grid_create").argtypes = [C.c_int, _dp, _dp, _up, _dp, C.c_double, C.c_double, C.c_double,
                                             C.c_uint, C.c_int, C.c_int]
            g("orc_grid_destroy").argtypes = [C.c_void_p]
            g("orc_grid_update").argtypes = [C.c_void_p, C.c_uint, _dp, _dp]
            g("orc_grid_update_deposit").argtypes = [C.c_void_p, C.c_uint, _dp]
            g("orc_grid_update_merge").argtypes = [C.c_void_p, C.c_uint, _dp, _dp]
            g("orc_grid_delta_io").argtypes = [C.c_void_p, C.c_int, C.c_int, _dp]
            g("orc_grid_compute_sigma").argtypes = [C.c_void_p, C.POINTER(_fp), C.c_uint, C.c_double, _dp]
            g("orc_grid_set_sigma_inv").argtypes = [C.c_void_p, _dp]
            g("orc_grid_get_sigma_inv").argtypes = [C.c_void_p, _dp]
            g("orc_grid_get").argtypes = [C.c_void_p, C.c_int, _dp]
            g("orc_grid_scalars").argtypes = [C.c_void_p, _dp]
            g("orc_grid_interpolate").restype = C.c_double
            g("orc_grid_interpolate").argtypes = [C.c_void_p, _dp, C.c_int]
            g("orc_grid_bin").restype = C.c_int
            g("orc_grid_bin").argtypes = [C.c_void_p, _dp]
            g("orc_grid_set_flags").argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint]
            g("orc_grid_reset_histogram").argtypes = [C.c_void_p]
            g("orc_grid_write").argtypes = [C.c_void_p, C.c_char_p, C.c_uint]
            g("orc_grid_read").restype = C.c_int
            g("orc_grid_read").argtypes = [C.c_void_p, C.c_char_p]
            g("orc_umbrella").restype = C.c_double
            g("orc_umbrella").argtypes = [C.c_int, C.c_int] + [C.c_double] * 6
            g("orc_wte_pe").restype = C.c_double
            g("orc_wte_pe").argtypes = [_fp, C.c_uint, C.c_double]
            g("orc_wte_scale").argtypes = [_fp, _fp, _fp, C.c_uint, C.c_uint, C.c_double, _dp]
            g("orc_aspect_value").restype = C.c_double
            g("orc_aspect_value").argtypes = [_dp, C.c_uint, C.c_uint]
            g("orc_aspect_virial").argtypes = [_dp, C.c_uint, C.c_uint, C.c_double, _dp]
            g("orc_density_value").restype = C.c_double
            g("orc_density_value").argtypes = [_dp, C.c_uint]
            g("orc_density_virial").argtypes = [_dp, C.c_uint, C.c_double, _dp]
            g("orc_fft3d").argtypes = [_dp, _dp, C.c_uint, C.c_uint, C.c_uint, C.c_int]
        _lib.orc_indexgrid_index.restype = C.c_uint
        _lib.orc_indexgrid_index.argtypes = [_up, C.c_int, _up]
        _lib.orc_indexgrid_coords.argtypes = [_up, C.c_int, C.c_uint, _up]
        _lib.orc_indexgrid_num.restype = C.c_uint
        _lib.orc_indexgrid_num.argtypes = [_up, C.c_int]
    return _lib


def _fn(name, prec):
    return getattr(lib(), "%s_%s" % (name, prec))


def box6(L, tilt=(0.0, 0.0, 0.0)):
    L = np.broadcast_to(np.asarray(L, dtype=np.float64), (3,))
    return np.ascontiguousarray(np.concatenate([L, np.asarray(tilt, dtype=np.float64)]))


def make_postype(pos, types=None):
    """(N,3) float positions + integer types -> HOOMD Scalar4 array (type id as raw bits in .w)."""
    pos = np.asarray(pos, dtype=np.float32)
    n = pos.shape[0]
    out = np.empty((n, 4), dtype=np.float32)
    out[:, :3] = pos
    t = np.zeros(n, dtype=np.int32) if types is None else np.asarray(types, dtype=np.int32)
    out[:, 3] = t.view(np.float32)
    return np.ascontiguousarray(out)


class Mesh:
    """OrderParameterMesh oracle (CPU path)."""

    def __init__(self, nx, ny, nz, mode, L, n_global, prec="f64", tilt=(0, 0, 0), literal_copysignf=True, literal_tilt_offset=True):
        """literal_copysignf=False: evaluate |x| exactly in assignTSCderiv (what a SINGLE_PRECISION build does);
        the double instance with this switch off is the tolerance target for forces (see metad_oracle.hpp).
        literal_tilt_offset=False: remove the constant the reference's in-cell offsets carry in a triclinic box."""
        self.prec = prec
        self.dims = (nx, ny, nz)
        self.M = nx * ny * nz
        mode = np.ascontiguousarray(mode, dtype=np.float64)
        self._box = box6(L, tilt)
        self.h = _fn("orc_mesh_create", prec)(nx, ny, nz, _d(mode), len(mode), _d(self._box), n_global)
        _fn("orc_mesh_set_literal_copysignf", prec)(self.h, int(literal_copysignf))
        _fn("orc_mesh_set_literal_tilt_offset", prec)(self.h, int(literal_tilt_offset))

    def __del__(self):
        try:
            _fn("orc_mesh_destroy", self.prec)(self.h)
        except Exception:
            pass

    def assign(self, postype):
        self._n = postype.shape[0]
        _fn("orc_mesh_assign", self.prec)(self.h, _f(postype), postype.shape[0])

    def update(self):
        _fn("orc_mesh_update", self.prec)(self.h)

    def cv(self):
        return _fn("orc_mesh_cv", self.prec)(self.h)

    def current_value(self, postype):
        self._n = postype.shape[0]
        return _fn("orc_mesh_current_value", self.prec)(self.h, _f(postype), postype.shape[0])

    def forces(self, postype, bias):
        out = np.empty((postype.shape[0], 4), dtype=np.float64)
        _fn("orc_mesh_forces", self.prec)(self.h, _f(postype), postype.shape[0], float(bias), _d(out))
        return out

    def _get(self, which, cplx):
        out = np.empty(self.M * (2 if cplx else 1), dtype=np.float64)
        _fn("orc_mesh_get", self.prec)(self.h, which, _d(out))
        nx, ny, nz = self.dims
        if cplx:
            return out.view(np.complex128).reshape(nz, ny, nx)
        return out.reshape(nz, ny, nx)

    mesh = property(lambda s: s._get(0, False))
    fourier = property(lambda s: s._get(1, True))
    fourier_G = property(lambda s: s._get(2, True))
    inv_re = property(lambda s: s._get(3, False))
    interp = property(lambda s: s._get(4, False))
    inv_im = property(lambda s: s._get(5, False))

    def cells(self):
        out = np.empty((self._n, 3), dtype=np.int32)
        _fn("orc_mesh_cells", self.prec)(self.h, out.ctypes.data_as(_ip), self._n)
        return out

    def mode_sq(self):
        return _fn("orc_mesh_mode_sq", self.prec)(self.h)

    def virial(self, table_d, kmin, kmax, bias, use_table=True):
        """computeVirial (k-space virial of the bias, OrderParameterMesh.cc:970-1050): external virial xx, xy, xz, yy, yz, zz."""
        t = np.ascontiguousarray(table_d, dtype=np.float64)
        out = np.empty(6, dtype=np.float64)
        _fn("orc_mesh_virial", self.prec)(self.h, _d(t), len(t), float(kmin), float(kmax), int(use_table), float(bias), _d(out))
        return out

    def qmax(self):
        out = np.empty(4, dtype=np.float64)
        _fn("orc_mesh_qmax", self.prec)(self.h, _d(out))
        return out


def lamellar_cv(postype, n_global, mode, lattice_vectors, L, prec="f64", tilt=(0, 0, 0)):
    mode = np.ascontiguousarray(mode, dtype=np.float64)
    lv = np.ascontiguousarray(lattice_vectors, dtype=np.int32).reshape(-1, 3)
    b = box6(L, tilt)
    modes = np.empty(2 * lv.shape[0], dtype=np.float64)
    cv = _fn("orc_lamellar_cv", prec)(_f(postype), postype.shape[0], n_global, _d(mode), len(mode),
                                      lv.ctypes.data_as(_ip), lv.shape[0], _d(b), _d(modes))
    return cv, modes.reshape(-1, 2)


def lamellar_forces(postype, n_global, mode, lattice_vectors, L, bias, prec="f64", tilt=(0, 0, 0)):
    mode = np.ascontiguousarray(mode, dtype=np.float64)
    lv = np.ascontiguousarray(lattice_vectors, dtype=np.int32).reshape(-1, 3)
    b = box6(L, tilt)
    out = np.empty((postype.shape[0], 4), dtype=np.float64)
    _fn("orc_lamellar_forces", prec)(_f(postype), postype.shape[0], n_global, _d(mode), len(mode),
                                     lv.ctypes.data_as(_ip), lv.shape[0], _d(b), float(bias), _d(out))
    return out


class Grid:
    """IntegratorMetaDynamics grid-mode oracle."""
    ARR = dict(grid=0, reweighted=1, weight=2, sigma_grid=3, hist=4, hist_gauss=5, hist_delta=6, grid_delta=7)

    def __init__(self, cv_min, cv_max, num_points, sigma, W=1.0, T_shift=1.0, T=1.0, stride=1, add_bias=True,
                 well_tempered=False, prec="f64"):
        self.prec = prec
        self.d = len(num_points)
        self.num_points = tuple(int(n) for n in num_points)
        a = lambda v: np.ascontiguousarray(v, dtype=np.float64)
        npts = np.ascontiguousarray(num_points, dtype=np.uint32)
        self.h = _fn("orc_grid_create", prec)(self.d, _d(a(cv_min)), _d(a(cv_max)), npts.ctypes.data_as(_up),
                                              _d(a(sigma)), W, T_shift, T, stride, int(add_bias), int(well_tempered))
        self.G = int(np.prod(self.num_points))

    def __del__(self):
        try:
            _fn("orc_grid_destroy", self.prec)(self.h)
        except Exception:
            pass

    def update(self, timestep, cv_vals):
        cur = np.ascontiguousarray(cv_vals, dtype=np.float64)
        out = np.empty(self.d, dtype=np.float64)
        _fn("orc_grid_update", self.prec)(self.h, int(timestep), _d(cur), _d(out))
        return out

    def get(self, name):
        out = np.empty(self.G, dtype=np.float64)
        _fn("orc_grid_get", self.prec)(self.h, self.ARR[name], _d(out))
        return out

    # the two halves of updateBiasPotential on either side of the multiple-walker all-reduce (IntegratorMetaDynamics.cc:392-410)
    def update_deposit(self, timestep, cv_vals):
        cur = np.ascontiguousarray(cv_vals, dtype=np.float64)
        _fn("orc_grid_update_deposit", self.prec)(self.h, int(timestep), _d(cur))

    def update_merge(self, timestep, cv_vals):
        cur = np.ascontiguousarray(cv_vals, dtype=np.float64)
        out = np.empty(self.d, dtype=np.float64)
        _fn("orc_grid_update_merge", self.prec)(self.h, int(timestep), _d(cur), _d(out))
        return out

    DELTAS = ("grid_delta", "sigma_grid_delta", "hist_delta", "hist_gauss_delta")

    def get_deltas(self):
        out = np.empty((4, self.G), dtype=np.float64)
        for k in range(4):
            _fn("orc_grid_delta_io", self.prec)(self.h, k, 0, _d(out[k]))
        return out

    def set_deltas(self, arr):
        arr = np.ascontiguousarray(arr, dtype=np.float64)
        for k in range(4):
            _fn("orc_grid_delta_io", self.prec)(self.h, k, 1, _d(arr[k]))

    def compute_sigma(self, forces, sigma_g):
        """computeSigma (adaptive Gaussians): forces = list of (N,4) float32 derivative arrays, None for a CV that cannot
        compute derivatives; returns and installs sigma_inv (d x d)."""
        arrs = [None if f is None else np.ascontiguousarray(f, dtype=np.float32) for f in forces]
        n = next(a.shape[0] for a in arrs if a is not None) if any(a is not None for a in arrs) else 0
        ptrs = (_fp * self.d)(*[(a.ctypes.data_as(_fp) if a is not None else _fp()) for a in arrs])
        out = np.empty(self.d * self.d, dtype=np.float64)
        _fn("orc_grid_compute_sigma", self.prec)(self.h, ptrs, n, float(sigma_g), _d(out))
        return out.reshape(self.d, self.d)

    def set_sigma_inv(self, m):
        m = np.ascontiguousarray(m, dtype=np.float64)
        _fn("orc_grid_set_sigma_inv", self.prec)(self.h, _d(m))

    def scalars(self):
        out = np.empty(4, dtype=np.float64)
        _fn("orc_grid_scalars", self.prec)(self.h, _d(out))
        return dict(bias_potential=out[0], reweight=out[1], num_gaussians=int(out[2]), out_of_bounds=int(out[3]))

    def interpolate(self, vals, reweight=False):
        v = np.ascontiguousarray(vals, dtype=np.float64)
        return _fn("orc_grid_interpolate", self.prec)(self.h, _d(v), int(reweight))

    def bin(self, vals):
        v = np.ascontiguousarray(vals, dtype=np.float64)
        return _fn("orc_grid_bin", self.prec)(self.h, _d(v))

    def set_flags(self, add_bias, well_tempered, stride):
        _fn("orc_grid_set_flags", self.prec)(self.h, int(add_bias), int(well_tempered), int(stride))

    def reset_histogram(self):
        _fn("orc_grid_reset_histogram", self.prec)(self.h)

    def write(self, filename, timestep):
        _fn("orc_grid_write", self.prec)(self.h, filename.encode(), int(timestep))

    def read(self, filename):
        if _fn("orc_grid_read", self.prec)(self.h, filename.encode()) != 0:
            raise RuntimeError("Error reading grid.")


UMBRELLA = dict(no_umbrella=0, linear=1, harmonic=2, wall=3, gaussian=4)


def umbrella_bias(kind, val, bias_in=0.0, cv0=0.0, kappa=1.0, width_flat=0.0, scale=1.0, prec="f64"):
    return _fn("orc_umbrella", prec)(0, UMBRELLA[kind], cv0, kappa, width_flat, scale, val, bias_in)


def umbrella_potential(kind, val, cv0=0.0, kappa=1.0, width_flat=0.0, scale=1.0, prec="f64"):
    return _fn("orc_umbrella", prec)(1, UMBRELLA[kind], cv0, kappa, width_flat, scale, val, 0.0)


def wte_pe(net_force4, ext_energy=0.0, prec="f64"):
    nf = np.ascontiguousarray(net_force4, dtype=np.float32)
    return _fn("orc_wte_pe", prec)(_f(nf), nf.shape[0], ext_energy)


def wte_scale(net_force4, net_torque4, net_virial, pitch, bias, ext_virial, prec="f64"):
    f = np.array(net_force4, dtype=np.float32, copy=True)
    t = np.array(net_torque4, dtype=np.float32, copy=True)
    v = np.array(net_virial, dtype=np.float32, copy=True)
    e = np.array(ext_virial, dtype=np.float64, copy=True)
    _fn("orc_wte_scale", prec)(_f(f), _f(t), _f(v), pitch, f.shape[0], bias, _d(e))
    return f, t, v, e


def aspect_value(L, d1, d2, prec="f64", tilt=(0, 0, 0)):
    return _fn("orc_aspect_value", prec)(_d(box6(L, tilt)), d1, d2)


def aspect_virial(L, d1, d2, bias, prec="f64", tilt=(0, 0, 0)):
    out = np.empty(6, dtype=np.float64)
    _fn("orc_aspect_virial", prec)(_d(box6(L, tilt)), d1, d2, bias, _d(out))
    return out


def density_value(L, n, prec="f64"):
    return _fn("orc_density_value", prec)(_d(box6(L)), n)


def density_virial(L, n, bias, prec="f64"):
    out = np.empty(6, dtype=np.float64)
    _fn("orc_density_virial", prec)(_d(box6(L)), n, bias, _d(out))
    return out


def fft3d(a, sign, prec="f64"):
    """a: complex array shaped (nz, ny, nx); unnormalised DFT, sign=-1 forward / +1 inverse."""
    a = np.ascontiguousarray(a, dtype=np.complex128)
    nz, ny, nx = a.shape
    out = np.empty_like(a)
    _fn("orc_fft3d", prec)(_d(a.view(np.float64)), _d(out.view(np.float64)), nx, ny, nz, sign)
    return out


def indexgrid_index(lengths, coords):
    l = np.ascontiguousarray(lengths, dtype=np.uint32)
    c = np.ascontiguousarray(coords, dtype=np.uint32)
    return lib().orc_indexgrid_index(l.ctypes.data_as(_up), len(l), c.ctypes.data_as(_up))


def indexgrid_coords(lengths, idx):
    l = np.ascontiguousarray(lengths, dtype=np.uint32)
    c = np.empty(len(l), dtype=np.uint32)
    lib().orc_indexgrid_coords(l.ctypes.data_as(_up), len(l), int(idx), c.ctypes.data_as(_up))
    return c


def indexgrid_num(lengths):
    l = np.ascontiguousarray(lengths, dtype=np.uint32)
    return int(lib().orc_indexgrid_num(l.ctypes.data_as(_up), len(l)))
