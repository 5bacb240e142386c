"""Thin Python owners of the C-ABI handles (include/metad_b200.h) operating on torch CUDA tensors.

torch is used for device memory and streams only; all arithmetic happens in libmetad_b200.so.
Particle arrays follow HOOMD's layout: float32 (N,4) postype with the type id as raw bits in column 3,
float32 (N,4) force with column 3 = per-particle energy (always 0 for collective variables).
"""
import ctypes as C

import numpy as np
import torch

from . import _abi
from ._abi import Box, check, lib


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _np_d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(C.POINTER(C.c_double))


def _require_postype(t):
    if not (t.is_cuda and t.dtype == torch.float32 and t.dim() == 2 and t.shape[1] == 4 and t.is_contiguous()):
        raise ValueError("postype must be a contiguous float32 CUDA tensor of shape (N,4)")


def make_postype(pos, types=None, device="cuda"):
    """(N,3) positions + integer type ids -> HOOMD Scalar4 layout on the device."""
    pos = np.asarray(pos, dtype=np.float32)
    out = np.empty((pos.shape[0], 4), dtype=np.float32)
    out[:, :3] = pos
    t = np.zeros(pos.shape[0], dtype=np.int32) if types is None else np.asarray(types, dtype=np.int32)
    out[:, 3] = t.view(np.float32)
    return torch.from_numpy(out).to(device)


class Lamellar:
    """LamellarOrderParameter device path (reference: LamellarOrderParameterGPU.cc:34-132)."""

    def __init__(self, mode, lattice_vectors):
        lv = np.ascontiguousarray(lattice_vectors, dtype=np.int32).reshape(-1, 3)
        m, mp = _np_d(mode)
        self.n_wave = lv.shape[0]
        self.h = C.c_void_p()
        check(lib.metad_lamellar_create(C.byref(self.h), self.n_wave, lv.ctypes.data_as(C.POINTER(C.c_int)), len(m), mp))
        self.modes = torch.zeros(2 * self.n_wave, dtype=torch.float64, device="cuda")
        self.cv = torch.zeros(1, dtype=torch.float64, device="cuda")

    def __del__(self):
        if getattr(self, "h", None):
            lib.metad_lamellar_destroy(self.h)
            self.h = None

    def compute_modes(self, postype, n_global, box, finalize=True):
        _require_postype(postype)
        check(lib.metad_lamellar_modes(self.h, _ptr(postype), postype.shape[0], int(n_global), C.byref(box),
                                       _ptr(self.modes), int(finalize), _ptr(self.cv), _stream()))
        return self.cv

    def finalize(self, n_global):
        check(lib.metad_lamellar_finalize(self.h, _ptr(self.modes), int(n_global), _ptr(self.cv), _stream()))
        return self.cv

    def forces(self, postype, n_global, box, bias, out=None):
        _require_postype(postype)
        if out is None:
            out = torch.empty_like(postype)
        check(lib.metad_lamellar_forces(self.h, _ptr(postype), _ptr(out), postype.shape[0], int(n_global), C.byref(box),
                                        _ptr(bias), _stream()))
        return out


class Mesh:
    """OrderParameterMesh device path (reference: OrderParameterMeshGPU.cc:157-506)."""

    def __init__(self, nx, ny, nz, mode):
        m, mp = _np_d(mode)
        self.dims = (int(nx), int(ny), int(nz))
        self.h = C.c_void_p()
        check(lib.metad_mesh_create(C.byref(self.h), int(nx), int(ny), int(nz), len(m), mp))
        self.cv = torch.zeros(1, dtype=torch.float64, device="cuda")
        self._n = 0

    def __del__(self):
        if getattr(self, "h", None):
            lib.metad_mesh_destroy(self.h)
            self.h = None

    def set(self, key, value):
        check(lib.metad_mesh_set(self.h, int(key), int(value)))

    def compute_cv(self, postype, n_global, box):
        _require_postype(postype)
        self._n = postype.shape[0]
        check(lib.metad_mesh_cv(self.h, _ptr(postype), postype.shape[0], int(n_global), C.byref(box), _ptr(self.cv), _stream()))
        return self.cv

    def forces(self, postype, n_global, box, bias, out=None):
        _require_postype(postype)
        if out is None:
            out = torch.empty_like(postype)
        check(lib.metad_mesh_forces(self.h, _ptr(postype), _ptr(out), postype.shape[0], int(n_global), C.byref(box),
                                    _ptr(bias), _stream()))
        return out

    def cells(self):
        out = np.empty((self._n, 3), dtype=np.int32)
        check(lib.metad_mesh_get(self.h, 0, out.ctypes.data_as(C.c_void_p)))
        return out

    def _mesh(self, which):
        nx, ny, nz = self.dims
        out = np.empty((nz, ny, nx), dtype=np.float32)
        check(lib.metad_mesh_get(self.h, which, out.ctypes.data_as(C.c_void_p)))
        return out

    def rho(self):
        return self._mesh(1)

    def inv(self):
        return self._mesh(2)

    STAGES = ("tile_order", "spread", "fft_x_fwd", "fft_y_fwd", "fft_z_fused", "fft_y_inv", "fft_x_inv", "gather")

    def timings(self):
        """Per-stage milliseconds of the last compute_cv + forces pair (profiling knob 2 must be on)."""
        out = np.empty(len(self.STAGES), dtype=np.float32)
        check(lib.metad_mesh_get(self.h, 4, out.ctypes.data_as(C.c_void_p)))
        return dict(zip(self.STAGES, out.tolist()))

    def mode_sq(self):
        out = np.empty(1, dtype=np.float64)
        check(lib.metad_mesh_get(self.h, 3, out.ctypes.data_as(C.c_void_p)))
        return float(out[0])

    def graph_launches(self):
        """CUDA-graph replays so far (knob 4)."""
        out = C.c_ulonglong(0)
        check(lib.metad_mesh_get(self.h, 7, C.byref(out)))
        return int(out.value)

    def stats(self):
        """Tile-order / fixed-point statistics of the last spread (synchronises)."""
        out = np.empty(6, dtype=np.float64)
        check(lib.metad_mesh_get(self.h, 5, out.ctypes.data_as(C.c_void_p)))
        return dict(rebuilds=int(out[0]), drifted=int(out[1]), outside_slab=int(out[2]), range_warnings=int(out[3]),
                    fx_scale=float(out[4]), calls_since_rebuild=int(out[5]))


    def set_table(self, dK, k_min, k_max, use_table=True):
        """Tabulated derivative of the convolution kernel (cv.mesh.set_kernel): enters the k-space virial only."""
        a, ap = _np_d(dK)
        check(lib.metad_mesh_set_table(self.h, ap, len(a), float(k_min), float(k_max), int(use_table)))

    def extras(self):
        """Epilogues of the last compute_cv with set(13, 1): dict(virial[6] (per unit bias factor), q_max[3], sq_max, flat)."""
        out = np.empty(12, dtype=np.float64)
        check(lib.metad_mesh_get(self.h, 10, out.ctypes.data_as(C.c_void_p)))
        return dict(virial=out[:6].copy(), q_max=out[6:9].copy(), sq_max=float(out[9]), flat=int(out[10]), amplitude=float(out[11]))

    def accumulator(self):
        """Width of the density accumulators: {"wide": 64-bit accumulation in use, "requested": what the cell loads of the
        last rebuild of the tile order ask for (1 = 32 bits suffice, 2 = 64 bits, 3 = beyond the range)}."""
        out = np.zeros(2, dtype=np.uint32)
        check(lib.metad_mesh_get(self.h, 9, out.ctypes.data_as(C.c_void_p)))
        return dict(wide=bool(out[0]), requested=int(out[1]))


class BiasGrid:
    """IntegratorMetaDynamics grid bias on the device (reference: IntegratorMetaDynamics.cc:363-451)."""
    ARR = dict(grid=(0, np.float64), reweighted=(1, np.float64), weight=(2, np.float64), sigma_grid=(3, np.float64),
               hist=(4, np.uint32), hist_gauss=(5, np.uint32), hist_delta=(6, np.uint32))

    def __init__(self, cv_min, cv_max, num_points, sigma, W=1.0, T_shift=1.0, T=1.0, stride=1, add_bias=True,
                 well_tempered=False):
        a, ap = _np_d(cv_min)
        b, bp = _np_d(cv_max)
        s, sp = _np_d(sigma)
        n = np.ascontiguousarray(num_points, dtype=np.uint32)
        self.d = len(n)
        self.h = C.c_void_p()
        check(lib.metad_grid_create(C.byref(self.h), self.d, ap, bp, n.ctypes.data_as(C.POINTER(C.c_uint)), sp, W, T_shift, T,
                                    int(stride), int(add_bias), int(well_tempered)))
        self.G = int(lib.metad_grid_num_elements(self.h))
        self.bias = torch.zeros(self.d, dtype=torch.float64, device="cuda")

    def __del__(self):
        if getattr(self, "h", None):
            lib.metad_grid_destroy(self.h)
            self.h = None

    def step(self, timestep, cv_values):
        """cv_values: float64 CUDA tensor with n_cv entries; returns the device tensor of dV/ds_i."""
        check(lib.metad_grid_step(self.h, int(timestep), _ptr(cv_values), _ptr(self.bias), _stream()))
        return self.bias

    # ---- multiple walkers: the step in two halves with an all-reduce of the delta arrays in between on deposit steps
    def is_deposit_step(self, timestep):
        return bool(lib.metad_grid_is_deposit_step(self.h, int(timestep)))

    def step_deposit(self, timestep, cv_values):
        check(lib.metad_grid_step_deposit(self.h, int(timestep), _ptr(cv_values), _stream()))

    def step_merge(self, timestep, cv_values):
        check(lib.metad_grid_step_merge(self.h, int(timestep), _ptr(cv_values), _ptr(self.bias), _stream()))
        return self.bias

    def deltas_export(self):
        """(double[2G] = grid_delta | sigma_grid_delta, int32[2G] = hist_delta | hist_gauss_delta) as device tensors."""
        if getattr(self, "_dd", None) is None:
            self._dd = torch.empty(2 * self.G, dtype=torch.float64, device="cuda")
            self._du = torch.empty(2 * self.G, dtype=torch.int32, device="cuda")
        check(lib.metad_grid_deltas_export(self.h, _ptr(self._dd), _ptr(self._du), _stream()))
        return self._dd, self._du

    def deltas_import(self, dd, du):
        check(lib.metad_grid_deltas_import(self.h, _ptr(dd), _ptr(du), _stream()))

    def step_walkers(self, timestep, cv_values, all_reduce_sum):
        """One step of a walker in multiple-walker mode (IntegratorMetaDynamics.cc:392-410): all_reduce_sum(tensor) sums a
        device tensor in place over the walkers' partition communicator."""
        self.step_deposit(timestep, cv_values)
        if self.is_deposit_step(timestep):
            dd, du = self.deltas_export()
            all_reduce_sum(dd)
            all_reduce_sum(du)
            self.deltas_import(dd, du)
        return self.step_merge(timestep, cv_values)

    def set_sigma_inv(self, m):
        a, ap = _np_d(np.asarray(m, dtype=np.float64).reshape(-1))
        check(lib.metad_grid_set_sigma_inv(self.h, ap))

    def get(self, name):
        which, dt = self.ARR[name]
        out = np.empty(self.G, dtype=dt)
        check(lib.metad_grid_download(self.h, which, out.ctypes.data_as(C.c_void_p)))
        return out

    def put(self, name, arr):
        which, dt = self.ARR[name]
        a = np.ascontiguousarray(arr, dtype=dt)
        check(lib.metad_grid_upload(self.h, which, a.ctypes.data_as(C.c_void_p)))

    def scalars(self):
        out = np.empty(4, dtype=np.float64)
        check(lib.metad_grid_scalars(self.h, out.ctypes.data_as(C.POINTER(C.c_double))))
        return dict(bias_potential=out[0], reweight=out[1], num_gaussians=int(out[2]), out_of_bounds=int(out[3]))

    def set_flags(self, add_bias, well_tempered, stride):
        check(lib.metad_grid_set_flags(self.h, int(add_bias), int(well_tempered), int(stride)))

    def set_num_gaussians(self, n):
        check(lib.metad_grid_set_num_gaussians(self.h, int(n)))

    def reset_histogram(self):
        check(lib.metad_grid_reset_histogram(self.h, _stream()))


UMBRELLA = dict(no_umbrella=0, linear=1, harmonic=2, wall=3, gaussian=4)


def umbrella_apply(kind, cv, bias_in=None, cv0=0.0, kappa=1.0, width_flat=0.0, scale=1.0, bias_out=None, energy_out=None):
    if bias_out is None:
        bias_out = torch.zeros(1, dtype=torch.float64, device="cuda")
    check(lib.metad_umbrella_apply(UMBRELLA[kind], cv0, kappa, width_flat, scale, _ptr(cv), _ptr(bias_in), _ptr(bias_out),
                                   _ptr(energy_out), _stream()))
    return bias_out


def force_dot(f_i, f_j, scale=1.0, out=None):
    """scale * sum_n f_i[n].xyz . f_j[n].xyz (computeSigma's sums of products of the CV derivatives), device double."""
    if out is None:
        out = torch.zeros(1, dtype=torch.float64, device="cuda")
    check(lib.metad_force_dot(_ptr(f_i), _ptr(f_j), f_i.shape[0], float(scale), _ptr(out), _stream()))
    return out


def wte_reduce(net_force, external_energy=0.0, out=None):
    if out is None:
        out = torch.zeros(1, dtype=torch.float64, device="cuda")
    check(lib.metad_wte_reduce(_ptr(net_force), net_force.shape[0], float(external_energy), _ptr(out), _stream()))
    return out


def wte_scale(net_force, net_torque, net_virial, pitch, bias):
    check(lib.metad_wte_scale(_ptr(net_force), _ptr(net_torque), _ptr(net_virial), int(pitch), net_force.shape[0], _ptr(bias),
                              _stream()))
