"""Generates tests/golden/golden.{json,npz}.

The reference ships no golden vectors and cannot be built here (HOOMD-blue absent), so these fixtures are
REGRESSION PINS of the CPU oracle (oracle/metad_oracle.hpp, "parity unpinned"), not outputs of the reference
itself.  They freeze the oracle's results on small seeded inputs so that later edits of the oracle, the
workload generators or the kernels cannot drift silently.  Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po                      # noqa: E402
from metadynamics_plugin_b200 import workloads          # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
gold, arrays = {}, {}

w = workloads.c1()
m = po.Mesh(*w["mesh"], w["mode"], w["L"], 1000, "f64", literal_copysignf=False)
cv = m.current_value(w["postype"])
u = w["umbrella"]
bias = po.umbrella_bias("harmonic", cv, 0.0, cv0=u["cv0"], kappa=u["kappa"])
gold["c1_cv"], gold["c1_bias"] = cv, bias
gold["c1_umbrella_energy"] = po.umbrella_potential("harmonic", cv, cv0=u["cv0"], kappa=u["kappa"])
arrays["c1_postype"] = w["postype"]
arrays["c1_forces"] = m.forces(w["postype"], bias)
m32 = po.Mesh(*w["mesh"], w["mode"], w["L"], 1000, "f32")
m32.assign(w["postype"])
arrays["c1_cells"] = m32.cells()

w2 = workloads.c2(N=4096)
cv2, modes = po.lamellar_cv(w2["postype"], 4096, w2["mode"], w2["lattice_vectors"], w2["L"])
gold["c2s_cv"] = cv2
arrays["c2s_modes"] = modes
arrays["c2s_forces"] = po.lamellar_forces(w2["postype"], 4096, w2["mode"], w2["lattice_vectors"], w2["L"], 0.75)

seq = [0.30, 0.31, 0.29, 0.35, 0.33, 0.36, 0.40, 0.38, 0.41, 0.45]
g = po.Grid(**w2["grid"], W=1.0, T_shift=7.0, T=1.0, stride=2, well_tempered=True)
biases = [float(g.update(t, [s])[0]) for t, s in enumerate(seq)]
gold["grid_cv_sequence"], gold["grid_bias_sequence"] = seq, biases
arrays["grid_after"] = g.get("grid")
arrays["grid_weight_after"] = g.get("weight")

with open(os.path.join(HERE, "golden.json"), "w") as fh:
    json.dump(gold, fh, indent=1)
np.savez_compressed(os.path.join(HERE, "golden.npz"), **arrays)
print("wrote", sorted(gold), sorted(arrays))
