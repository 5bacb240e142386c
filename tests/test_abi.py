"""The C-ABI library loads on a CPU-only box and exports exactly what include/metad_b200.h declares."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "metad_b200.h")


def header_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(metad_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from metadynamics_plugin_b200 import _abi
    declared = header_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(_abi.lib, name), "libmetad_b200.so does not export %s" % name
    assert sorted(_abi.SIGNATURES) == declared                     # the ctypes binding covers the header one to one
    exported = subprocess.check_output(["nm", "-D", "--defined-only", _abi.LIB_PATH]).decode()
    extra = [s for s in re.findall(r" T (metad_[a-z0-9_]+)", exported) if s not in declared]
    assert not extra, "exported but not declared in the header: %s" % extra


def test_no_cpu_fallback_error_is_loud():
    """Without a CUDA device a compute entry point must fail with METAD_ERR_CUDA, not silently succeed."""
    import ctypes as C
    import numpy as np
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    from metadynamics_plugin_b200 import _abi
    assert _abi.lib.metad_version() >= 100
    h = C.c_void_p()
    mode = np.array([1.0])
    rc = _abi.lib.metad_mesh_create(C.byref(h), 32, 32, 32, 1, mode.ctypes.data_as(C.POINTER(C.c_double)))
    assert rc == -2 and "CUDA" in _abi.last_error()
    with pytest.raises(_abi.MetadError):
        _abi.check(rc)
    # argument validation happens before any CUDA call
    rc = _abi.lib.metad_mesh_create(C.byref(h), 2048, 32, 32, 1, mode.ctypes.data_as(C.POINTER(C.c_double)))
    assert rc == -3 and "1024" in _abi.last_error()
    rc = _abi.lib.metad_mesh_slab_create(C.byref(h), 96, 32, 32, 2, 0, 1, mode.ctypes.data_as(C.POINTER(C.c_double)))
    assert rc == -3 and "power of two" in _abi.last_error()
    # a mesh that is not a power of two is a valid plan (general path): without a device its creation fails on the first CUDA call
    rc = _abi.lib.metad_mesh_create(C.byref(h), 48, 32, 32, 1, mode.ctypes.data_as(C.POINTER(C.c_double)))
    assert rc == -2 and "CUDA" in _abi.last_error()


def test_product_package_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under metadynamics_plugin_b200/ may reference it."""
    pkg = os.path.join(ROOT, "metadynamics_plugin_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "pyoracle" not in text and "liboracle" not in text and "metad_oracle" not in text, f
