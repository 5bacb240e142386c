"""Parity sweep on the GPU: CV / Re IFFT(G) / force errors of the mesh path against the double oracle for a list of mesh
shapes and particle densities (diagnostic companion of tests/test_gpu_parity.py::test_mesh_every_fft_length).
The oracle is the checker here, as in the tests."""
import sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from metadynamics_plugin_b200 import ops
from oracle import pyoracle as po

cases = [(32, 256, 16), (32, 16, 256), (512, 16, 16), (1024, 16, 16), (32, 512, 16), (32, 16, 512), (64, 128, 32), (256, 32, 128)]
for dims in cases:
    M = int(np.prod(dims))
    for N in (30000, min(M, 400000)):
        Lf = np.asarray(dims, float) * 0.31
        rng = np.random.default_rng(sum(dims))
        pos = ((rng.random((N, 3)) - 0.5) * Lf).astype(np.float32)
        types = rng.integers(0, 2, N).astype(np.int32)
        modes = (1.0, -0.7)
        d_pt = ops.make_postype(pos, types)
        h_pt = po.make_postype(pos, types)
        box = ops.Box.make(Lf)
        mesh = ops.Mesh(*dims, modes)
        cv = mesh.compute_cv(d_pt, N, box).cpu().item()
        m = po.Mesh(*dims, modes, Lf, N, "f64", literal_copysignf=False)
        cvo = m.current_value(h_pt)
        inv, inv_o = np.asarray(mesh.inv(), dtype=np.float64), m.inv_re
        inv -= inv.mean(); inv_o = inv_o - inv_o.mean()
        f = mesh.forces(d_pt, N, box, torch.tensor([0.77], dtype=torch.float64, device="cuda")).cpu().numpy()
        fo = m.forces(h_pt, 0.77)
        print(json.dumps(dict(dims=dims, N=N, cv_rel=abs(cv / cvo - 1), inv_rel=float(np.abs(inv - inv_o).max() / np.abs(inv_o).max()),
                              f_rel=float(np.abs(f - fo).max() / np.abs(fo).max()), fmax=float(np.abs(fo).max()),
                              inv_max=float(np.abs(inv_o).max()))), flush=True)
