"""Where does the force error of a long-y mesh come from?  Error spectrum of Re IFFT(G) (device minus double oracle)."""
import sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from metadynamics_plugin_b200 import ops
from oracle import pyoracle as po

for dims in [(32, 512, 16), (32, 16, 512), (1024, 16, 16)]:
    N = 262144
    Lf = np.asarray(dims, float) * 0.31
    rng = np.random.default_rng(sum(dims))
    pos = ((rng.random((N, 3)) - 0.5) * Lf).astype(np.float32)
    types = rng.integers(0, 2, N).astype(np.int32)
    modes = (1.0, -0.7)
    d_pt = ops.make_postype(pos, types)
    h_pt = po.make_postype(pos, types)
    box = ops.Box.make(Lf)
    mesh = ops.Mesh(*dims, modes)
    mesh.set(1, 1)
    cv = mesh.compute_cv(d_pt, N, box).cpu().item()
    m = po.Mesh(*dims, modes, Lf, N, "f64", literal_copysignf=False)
    cvo = m.current_value(h_pt)
    rho_err = np.asarray(mesh.rho(), dtype=np.float64) - m.mesh
    inv, inv_o = np.asarray(mesh.inv(), dtype=np.float64), m.inv_re
    inv -= inv.mean(); inv_o = inv_o - inv_o.mean()
    err = inv - inv_o
    E = np.abs(np.fft.fftn(err)); S = np.abs(np.fft.fftn(inv_o))
    idx = np.argsort(E.ravel())[::-1][:8]
    top = [(tuple(int(v) for v in np.unravel_index(i, E.shape)), float(E.ravel()[i]), float(S.ravel()[i])) for i in idx]
    # force-like measure: central differences along each axis (axis order of the arrays: z, y, x)
    def grad_err(ax):
        return float(np.abs(np.roll(err, -1, ax) - np.roll(err, 1, ax)).max() / np.abs(np.roll(inv_o, -1, ax) - np.roll(inv_o, 1, ax)).max())
    f = mesh.forces(d_pt, N, box, torch.tensor([0.77], dtype=torch.float64, device="cuda")).cpu().numpy()
    fo = m.forces(h_pt, 0.77)
    fe = np.abs(f - fo)[:, :3].max(0) / np.abs(fo).max()
    print(json.dumps(dict(dims=dims, rho_err_max=float(np.abs(rho_err).max()), inv_rel=float(np.abs(err).max() / np.abs(inv_o).max()),
                          grad_err_zyx=[grad_err(0), grad_err(1), grad_err(2)], f_err_xyz=[float(v) for v in fe],
                          top_modes_zyx_err_signal=top, E_rms=float(np.sqrt((E ** 2).mean())), S_rms=float(np.sqrt((S ** 2).mean())))), flush=True)
