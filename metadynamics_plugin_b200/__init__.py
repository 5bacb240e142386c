"""metadynamics_plugin_b200 -- B200-native (sm_100a) CV + bias-force hot path of the HOOMD-blue metadynamics plugin.

Layers (bottom up):
  csrc/                 hand-written CUDA kernels + the C ABI (include/metad_b200.h) -> libmetad_b200.so
  _abi.py               ctypes binding of that C ABI (raises if the library is missing: no CPU fallback)
  ops.py                thin owners of the opaque C handles working on torch CUDA tensors (device memory, streams)
  host/                 C++ host classes with the reference's operator surface + pybind11 module `_metadynamics`
  cv.py, integrate.py   the reference's Python API (cv.lamellar / cv.mesh / ... / integrate.mode_metadynamics)
"""
__version__ = "0.1.0"
