import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure): built on demand with its committed Makefile."""
    from oracle import pyoracle
    pyoracle.build()
    pyoracle.lib()
    return pyoracle


NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def build_emul(name):
    """Build one of the host-only CUDA-free emulation programs in tests/cpu_emul with nvcc (no GPU needed)."""
    src = os.path.join(ROOT, "tests", "cpu_emul", name + ".cu")
    out_dir = os.path.join(ROOT, "tests", "cpu_emul", "_build")
    os.makedirs(out_dir, exist_ok=True)
    exe = os.path.join(out_dir, name)
    deps = [src, os.path.join(ROOT, "tests", "cpu_emul", "emul_fft.h")] + [
        os.path.join(ROOT, "metadynamics_plugin_b200", "csrc", f)
        for f in ("mesh_fft.cuh", "mesh_fft_kernels.cuh", "mesh_kernels.cuh", "mesh_general.cuh", "common.cuh")]
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(d) for d in deps):
        if not os.path.exists(NVCC):
            pytest.skip("nvcc not available to build the CPU emulation harness")
        subprocess.check_call([NVCC, "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler",
                               "-ffp-contract=off,-Wno-unknown-pragmas", "-o", exe, src])
    return exe


def triclinic_case(N, L, tilt, ntypes, seed, faces=True):
    """Particles inside a sheared box (HOOMD convention: r = lo + f L, x += xy y + xz z, y += yz z, all in float like a
    single-precision trajectory), some of them a few ulps around cell faces of the fractional mesh."""
    rng = np.random.default_rng(seed)
    Lf = np.asarray(L, dtype=np.float64)
    f = rng.random((N, 3))
    if faces:
        k = min(N // 4, 400)
        f[:k] = np.round(f[:k] * 16) / 16 + rng.integers(-3, 4, (k, 3)) * 2.0 ** -24
        f = np.clip(f, 2.0 ** -20, 1.0 - 2.0 ** -20)      # inside the box after the rounding to float (HOOMD wraps the rest)
    v = (f - 0.5) * Lf
    xy, xz, yz = tilt
    v[:, 0] += xy * v[:, 1] + xz * v[:, 2]
    v[:, 1] += yz * v[:, 2]
    return v.astype(np.float32), rng.integers(0, ntypes, N).astype(np.int32)
