"""ctypes binding of the C ABI in include/metad_b200.h (libmetad_b200.so, sm_100a only).

There is no CPU fallback: importing this module without the built library raises, and every call
on a machine without a CUDA device fails with the library's own error text.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmetad_b200.so")

METAD_OK = 0
ERR_NAMES = {-1: "METAD_ERR_INVALID", -2: "METAD_ERR_CUDA", -3: "METAD_ERR_UNSUPPORTED", -4: "METAD_ERR_STATE"}


class MetadError(RuntimeError):
    def __init__(self, code, text):
        super().__init__("%s: %s" % (ERR_NAMES.get(code, str(code)), text))
        self.code = code


class Box(C.Structure):
    """metad_box: box lengths and tilt factors (HOOMD BoxDim flattened)."""
    _fields_ = [("L", C.c_double * 3), ("tilt", C.c_double * 3)]

    @classmethod
    def make(cls, L, tilt=(0.0, 0.0, 0.0)):
        try:
            Lx, Ly, Lz = (float(v) for v in L)
        except TypeError:
            Lx = Ly = Lz = float(L)
        b = cls()
        b.L[0], b.L[1], b.L[2] = Lx, Ly, Lz
        b.tilt[0], b.tilt[1], b.tilt[2] = (float(t) for t in tilt)
        return b


if not os.path.exists(LIB_PATH):
    raise ImportError(
        "libmetad_b200.so is not built (expected at %s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
        "or `make -C metadynamics_plugin_b200/csrc`. There is no CPU fallback." % LIB_PATH)

lib = C.CDLL(LIB_PATH)

_vp, _dp, _ip, _up = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_uint)
_boxp = C.POINTER(Box)

# name -> (restype, argtypes); mirrors include/metad_b200.h one to one (tests check the list against the header)
SIGNATURES = {
    "metad_version": (C.c_int, []),
    "metad_last_error": (C.c_char_p, []),
    "metad_lamellar_create": (C.c_int, [C.POINTER(_vp), C.c_int, _ip, C.c_int, _dp]),
    "metad_lamellar_destroy": (C.c_int, [_vp]),
    "metad_lamellar_modes": (C.c_int, [_vp, _vp, C.c_uint, C.c_uint, _boxp, _vp, C.c_int, _vp, _vp]),
    "metad_lamellar_finalize": (C.c_int, [_vp, _vp, C.c_uint, _vp, _vp]),
    "metad_lamellar_forces": (C.c_int, [_vp, _vp, _vp, C.c_uint, C.c_uint, _boxp, _vp, _vp]),
    "metad_mesh_create": (C.c_int, [C.POINTER(_vp), C.c_uint, C.c_uint, C.c_uint, C.c_int, _dp]),
    "metad_mesh_destroy": (C.c_int, [_vp]),
    "metad_mesh_set_table": (C.c_int, [_vp, C.POINTER(C.c_double), C.c_uint, C.c_double, C.c_double, C.c_int]),
    "metad_mesh_cv": (C.c_int, [_vp, _vp, C.c_uint, C.c_uint, _boxp, _vp, _vp]),
    "metad_mesh_forces": (C.c_int, [_vp, _vp, _vp, C.c_uint, C.c_uint, _boxp, _vp, _vp]),
    "metad_mesh_slab_create": (C.c_int, [C.POINTER(_vp), C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_int, _dp]),
    "metad_mesh_slab_spread": (C.c_int, [_vp, _vp, C.c_uint, _boxp, _vp, _vp, _vp]),
    "metad_mesh_slab_fft_x": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "metad_mesh_slab_fft_yz": (C.c_int, [_vp, _vp, _vp, C.c_uint, _vp, _vp]),
    "metad_mesh_slab_fft_x_inv": (C.c_int, [_vp, _vp, _vp, _vp]),
    "metad_mesh_slab_forces": (C.c_int, [_vp, _vp, _vp, _vp, C.c_uint, C.c_uint, _boxp, _vp, _vp]),
    "metad_mesh_slab_p2p_arena": (C.c_int, [_vp, _vp, C.POINTER(C.c_ulonglong)]),
    "metad_mesh_slab_p2p_connect": (C.c_int, [_vp, _vp]),
    "metad_mesh_slab_p2p_connect_local": (C.c_int, [_vp, C.POINTER(_vp)]),
    "metad_mesh_slab_p2p_cv": (C.c_int, [_vp, _vp, C.c_uint, C.c_uint, _boxp, _vp, C.c_int, _vp]),
    "metad_mesh_slab_p2p_forces": (C.c_int, [_vp, _vp, _vp, C.c_uint, C.c_uint, _boxp, _vp, _vp]),
    "metad_mesh_get": (C.c_int, [_vp, C.c_int, _vp]),
    "metad_mesh_set": (C.c_int, [_vp, C.c_int, C.c_long]),
    "metad_grid_create": (C.c_int, [C.POINTER(_vp), C.c_int, _dp, _dp, _up, _dp, C.c_double, C.c_double, C.c_double,
                                    C.c_uint, C.c_int, C.c_int]),
    "metad_grid_destroy": (C.c_int, [_vp]),
    "metad_grid_step": (C.c_int, [_vp, C.c_uint, _vp, _vp, _vp]),
    "metad_grid_step_deposit": (C.c_int, [_vp, C.c_uint, _vp, _vp]),
    "metad_grid_step_merge": (C.c_int, [_vp, C.c_uint, _vp, _vp, _vp]),
    "metad_grid_is_deposit_step": (C.c_int, [_vp, C.c_uint]),
    "metad_grid_deltas_export": (C.c_int, [_vp, _vp, _vp, _vp]),
    "metad_grid_deltas_import": (C.c_int, [_vp, _vp, _vp, _vp]),
    "metad_grid_set_sigma_inv": (C.c_int, [_vp, C.POINTER(C.c_double)]),
    "metad_set_double": (C.c_int, [_vp, C.c_double, _vp]),
    "metad_force_dot": (C.c_int, [_vp, _vp, C.c_uint, C.c_double, _vp, _vp]),
    "metad_grid_set_flags": (C.c_int, [_vp, C.c_int, C.c_int, C.c_uint]),
    "metad_grid_reset_histogram": (C.c_int, [_vp, _vp]),
    "metad_grid_download": (C.c_int, [_vp, C.c_int, _vp]),
    "metad_grid_upload": (C.c_int, [_vp, C.c_int, _vp]),
    "metad_grid_scalars": (C.c_int, [_vp, _dp]),
    "metad_grid_set_num_gaussians": (C.c_int, [_vp, C.c_uint]),
    "metad_grid_num_elements": (C.c_uint, [_vp]),
    "metad_umbrella_apply": (C.c_int, [C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, _vp, _vp, _vp, _vp, _vp]),
    "metad_peer_create": (C.c_int, [C.POINTER(_vp), C.c_uint, C.c_uint]),
    "metad_peer_destroy": (C.c_int, [_vp]),
    "metad_peer_handle": (C.c_int, [_vp, _vp]),
    "metad_peer_connect": (C.c_int, [_vp, _vp]),
    "metad_peer_connect_local": (C.c_int, [_vp, _vp]),
    "metad_peer_allreduce_sum": (C.c_int, [_vp, _vp, C.c_uint, C.c_int, _vp]),
    "metad_peer_status": (C.c_int, [_vp, C.POINTER(C.c_uint)]),
    "metad_wte_reduce": (C.c_int, [_vp, C.c_uint, C.c_double, _vp, _vp]),
    "metad_wte_scale": (C.c_int, [_vp, _vp, _vp, C.c_uint, C.c_uint, _vp, _vp]),
    "metad_accumulate_force": (C.c_int, [_vp, _vp, C.c_uint, C.c_int, _vp]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _f = getattr(lib, _name)      # AttributeError here = the library does not export what the header declares
    _f.restype = _res
    _f.argtypes = _args


def last_error():
    return (lib.metad_last_error() or b"").decode()


def check(rc):
    if rc != METAD_OK:
        raise MetadError(rc, last_error())
    return rc
