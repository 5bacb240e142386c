import os
import subprocess
import sys

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
        for f in ("mesh_fft.cuh", "mesh_fft_kernels.cuh", "mesh_kernels.cuh", "common.cuh")]
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(d) for d in deps):
        if not os.path.exists(NVCC):
            pytest.skip("nvcc not available to build the CPU emulation harness")
        subprocess.check_call([NVCC, "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler",
                               "-ffp-contract=off,-Wno-unknown-pragmas", "-o", exe, src])
    return exe
