"""Per-stage timings of the rows added for SURVEY 8 (a3): cv.mesh in a triclinic box (tiled path) and on mesh sizes that are
not powers of two (general path), next to the orthorhombic tiled path on the same particles.  Writes one JSON object;
run on the GPU box:  python tools/a3_timing.py > gpurun_out/a3_timing.json"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from metadynamics_plugin_b200 import ops  # noqa: E402


def particles(N, L, tilt, nmesh, seed):
    """Uniform particles inside the (sheared) box, sorted by mesh cell of their fractional coordinates."""
    rng = np.random.default_rng(seed)
    f = rng.random((N, 3))
    cell = (f * nmesh).astype(np.int64)
    f = f[np.argsort(cell[:, 0] + nmesh * (cell[:, 1] + nmesh * cell[:, 2]), kind="stable")]
    v = (np.clip(f, 1e-6, 1 - 1e-6) - 0.5) * L
    v[:, 0] += tilt[0] * v[:, 1] + tilt[1] * v[:, 2]
    v[:, 1] += tilt[2] * v[:, 2]
    return v.astype(np.float32)


def run(dims, pos, L, tilt, knob16=1, force_general=False, reps=6):
    if force_general:
        os.environ["METAD_MESH_GENERAL"] = "1"
    else:
        os.environ.pop("METAD_MESH_GENERAL", None)
    N = pos.shape[0]
    pt = ops.make_postype(pos, np.zeros(N, np.int32))
    box = ops.Box.make([L] * 3, tilt)
    mesh = ops.Mesh(*dims, [1.0])
    mesh.set(16, knob16)
    mesh.set(2, 1)
    bias = torch.tensor([0.5], dtype=torch.float64, device="cuda")
    acc = None
    for i in range(reps):
        mesh.compute_cv(pt, N, box)
        mesh.forces(pt, N, box, bias)
        if i >= 2:
            t = mesh.timings()
            acc = t if acc is None else {k: acc[k] + t[k] for k in t}
    out = {k: round(v / (reps - 2), 5) for k, v in acc.items()}
    out["total"] = round(sum(out.values()), 5)
    return out


def main():
    N, nmesh = 1 << 20, 128
    L = float(N) ** (1.0 / 3.0)
    tilt = (0.2, -0.1, 0.15)
    res = dict(N=N, box_L=L, tilt=tilt, unit="ms per stage, mean of 4 calls after 2 warm-up calls (events inside the library)")
    p0 = particles(N, L, (0.0, 0.0, 0.0), nmesh, 1)
    p1 = particles(N, L, tilt, nmesh, 1)
    res["tiled_128_orthorhombic"] = run((nmesh,) * 3, p0, L, (0.0, 0.0, 0.0))
    res["tiled_128_triclinic_literal"] = run((nmesh,) * 3, p1, L, tilt, 1)
    res["tiled_128_triclinic_corrected"] = run((nmesh,) * 3, p1, L, tilt, 0)
    res["general_128_orthorhombic"] = run((nmesh,) * 3, p0, L, (0.0, 0.0, 0.0), force_general=True)
    res["general_96_orthorhombic"] = run((96,) * 3, p0, L, (0.0, 0.0, 0.0))
    res["general_100_triclinic_corrected"] = run((100,) * 3, p1, L, tilt, 0)
    res["general_127_prime_orthorhombic"] = run((127,) * 3, p0, L, (0.0, 0.0, 0.0))
    print(json.dumps(res))


if __name__ == "__main__":
    main()
