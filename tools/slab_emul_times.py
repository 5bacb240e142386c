#!/usr/bin/env python
"""All ranks of a P-way z-slab decomposition of one workload on ONE GPU (the launch order replaces the flag barriers,
peer stores land in local memory): the per-rank kernels at their multi-GPU problem sizes, for ncu launch lists
(`ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none ... python tools/slab_emul_times.py 8`).

    python tools/slab_emul_times.py <ranks> [workload] [steps]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from metadynamics_plugin_b200 import ops, sharded, workloads       # noqa: E402

P = int(sys.argv[1])
name = sys.argv[2] if len(sys.argv) > 2 else "C4"
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
w = getattr(workloads, name.lower())()
owner = sharded.slab_of(w["postype"][:, 2], w["L"], w["mesh"][2], P)
pts = [torch.from_numpy(np.ascontiguousarray(w["postype"][owner == r])).cuda() for r in range(P)]
ranks = [sharded.MeshSlabRank(*w["mesh"], P, r, w["mode"]) for r in range(P)]
sharded.connect_local(ranks)
box = ops.Box.make(w["L"])
bias = torch.tensor([0.7], dtype=torch.float64, device="cuda")
N = w["postype"].shape[0]
for _ in range(steps):
    cvs, forces = sharded.mesh_slab_p2p_step_local(ranks, pts, N, box, bias)
torch.cuda.synchronize()
single = ops.Mesh(*w["mesh"], w["mode"])
cv1 = single.compute_cv(torch.from_numpy(w["postype"]).cuda(), N, box).cpu().item()
print("cv sharded", cvs[0].cpu().item(), "single", cv1)
