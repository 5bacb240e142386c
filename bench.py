#!/usr/bin/env python
"""bench.py -- CV + bias-force step benchmark (BASELINE.json metric), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C4|C3|C2|C5|C1] [--impl ours|reference]

A "step" = one pass of the hot path over one batch of synthetic particles:
    mesh workloads      metad_mesh_cv -> metad_grid_step (1-D grid bias, deposit every `stride`) -> metad_mesh_forces
    lamellar workloads  metad_lamellar_modes -> metad_grid_step -> metad_lamellar_forces
Rank 0 prints ONE JSON line (see DESIGN.md "Measurement" for every field).  `value` is measured with the inputs
resident in HBM; `e2e` repeats the step through the same calls with HOST (pinned) buffers, host->device copy of the
positions and device->host copy of the forces and the CV inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

STAGE_BYTES_DOC = "algorithmic bytes per launch: see DESIGN.md 'Kernels and rooflines'"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


PARTICLE_ORDER = "sorted"   # --order: C3 / C4 input order, cell-sorted (an MD engine's spatial sort; headline) or random (stress)


def make_workload(name, shard=None, order=None):
    from metadynamics_plugin_b200 import workloads
    order = order or PARTICLE_ORDER
    if name in ("C3", "C4") and order in ("random", "sfc"):
        w = {"C3": workloads.c3, "C4": workloads.c4}[name](sort=False if order == "random" else "sfc")
        w["order"] = ("random particle order (stress case, SURVEY 8d C3 ii)" if order == "random" else
                      "Morton order on 4^3-cell blocks, random inside a block (an MD engine's SFC sort)")
        return w
    return {"C1": workloads.c1, "C2": workloads.c2, "C3": workloads.c3, "C4": workloads.c4, "C5": workloads.c5, "WTE": workloads.wte}[name]()


# ---------------------------------------------------------------------------------------------------- ours
MERGE_PUSH = False      # --merge-push: peer-memory mode with halo push and barrier in one launch (measured slower)
NO_PDL = False          # --no-pdl: A/B switch of programmatic dependent launch
MESH_KNOBS = []         # --mesh-knob K=V: metad_mesh_set(plan, K, V) before the first step (A/B experiments)
TILE_ORDER = None       # --tile-order: A/B switch of the order inside a tile (library default when None)


class MeshStep:
    """mesh CV -> 1-D grid bias -> forces, everything device-resident."""
    # spread, x/y fwd, z fused (+plane0), y/x inv, grid step, gather (no memset / copy nodes); a rebuild of the tile order adds bin,
    # 3 scan, place, bank order, scale
    launches_per_step = 8
    launches_per_rebuild = 7

    def __init__(self, w, ops, torch, calibrate=True, period=32):
        self.ops, self.torch, self.w = ops, torch, w
        self.N = w["postype"].shape[0]
        self.box = ops.Box.make(w["L"])
        self.mesh = ops.Mesh(*w["mesh"], w["mode"])
        self.period = period
        self.mesh.set(0, period)
        self.mesh.set(4, 1)             # CUDA-graph replay of the per-call kernel sequence
        if TILE_ORDER is not None:
            self.mesh.set(6, {"bank": 1, "layer": 0}[TILE_ORDER])
        if NO_PDL:
            self.mesh.set(7, 0)
        for k, v in MESH_KNOBS:
            self.mesh.set(k, v)
        self.d_pt = torch.from_numpy(w["postype"]).cuda()
        self.d_force = torch.empty_like(self.d_pt)
        self.t = 0
        # 1-D grid spanning the observed CV +-50 %, 400 points
        cv = self.mesh.compute_cv(self.d_pt, self.N, self.box).cpu().item()
        self.cv0 = cv
        lo, hi = (0.5 * cv, 1.5 * cv) if cv > 0 else (1.5 * cv, 0.5 * cv)
        self.grid = ops.BiasGrid([lo], [hi], [400], [0.05 * abs(cv)], W=1e-3 * abs(cv), T_shift=7.0, T=1.0,
                                 stride=w.get("stride", 100), well_tempered=True)

    def step(self):
        cv = self.mesh.compute_cv(self.d_pt, self.N, self.box)
        bias = self.grid.step(self.t, cv)
        self.mesh.forces(self.d_pt, self.N, self.box, bias, out=self.d_force)
        self.t += 1

    def algorithmic_bytes(self):
        M = int(np.prod(self.w["mesh"]))
        return 48 * self.N + 48 * M

    def stage_bytes(self):
        N, M = self.N, int(np.prod(self.w["mesh"]))
        return {"tile_order": 0, "spread": 16 * N + 4 * M, "fft_x_fwd": 8 * M, "fft_y_fwd": 8 * M,
                "fft_z_fused": 8 * M, "fft_y_inv": 8 * M, "fft_x_inv": 8 * M, "gather": 32 * N + 4 * M}


class MeshSlabStep(MeshStep):
    """The same step with the mesh sharded into z slabs over the ranks (sharded.MeshSlab: NCCL all-to-all transposes,
    halo exchanges and two tiny all-reduces per step); the bias grid is replicated.  Strong scaling: N is the global
    particle number, every rank owns the particles of its slab."""
    # our kernels per step.  peer memory: spread, 2 halo pushes, 4 barriers, x/y forward, fused z, y/x inverse, grid step, gather;
    # staged (NCCL): the same 8 compute kernels as on one GPU (the collectives are NCCL's kernels, not counted)
    launches_per_step = 14

    def __init__(self, w, ops, torch, comm, period=32, mode="p2p", sync="barrier"):
        from metadynamics_plugin_b200 import sharded
        self.comm_mode = mode
        self.ops, self.torch, self.w = ops, torch, w
        self.N_global = w["postype"].shape[0]
        owner = sharded.slab_of(w["postype"][:, 2], w["L"], w["mesh"][2], comm.size)
        local = np.ascontiguousarray(w["postype"][owner == comm.rank])
        self.N = local.shape[0]
        self.h_local = local
        self.box = ops.Box.make(w["L"])
        # reference result of the same step with library collectives (NCCL): the peer-memory path must reproduce it
        nccl = sharded.MeshSlab(comm, *w["mesh"], w["mode"])
        self.d_pt = torch.from_numpy(local).cuda()
        cv_nccl = nccl.compute_cv(self.d_pt, self.N_global, self.box).cpu().item()
        if mode == "p2p":
            try:
                self.slab = sharded.MeshSlabP2P(comm, *w["mesh"], w["mode"])
            except RuntimeError as e:       # raised on every rank together (see MeshSlabP2P): fall back to the staged path
                if comm.rank == 0:
                    print("bench: %s -- falling back to --comm nccl" % e, file=sys.stderr)
                mode = self.comm_mode = "nccl"
        if mode == "p2p":
            cv_p2p = self.slab.compute_cv(self.d_pt, self.N_global, self.box).cpu().item()
            self.cv_check = abs(cv_p2p / cv_nccl - 1.0)
            st = self.slab.status()
            if st["barrier_timeout"] or not self.cv_check < 1e-9:
                raise RuntimeError("peer-memory path disagrees with the NCCL path: cv %r vs %r, status %r" % (cv_p2p, cv_nccl, st))
            del nccl
        else:
            self.slab = nccl
            self.cv_check = 0.0
        self.mesh = self.slab.r
        self.launches_per_step = 14 if mode == "p2p" else 8
        self.period = period
        self.mesh.set(0, period)
        if NO_PDL:
            self.mesh.set(7, 0)
        if MERGE_PUSH:
            self.mesh.set(8, 1)
        for k, v in MESH_KNOBS:
            self.mesh.set(k, v)
        if mode == "p2p":
            self.mesh.set(5, 1 if sync == "fused" else 0)
            self.mesh.set(4, 1)         # the whole sharded step (incl. the inter-rank waits) replays from one CUDA graph
        self.d_force = torch.empty_like(self.d_pt)
        self.t = 0
        cv = self.slab.compute_cv(self.d_pt, self.N_global, self.box).cpu().item()
        lo, hi = (0.5 * cv, 1.5 * cv) if cv > 0 else (1.5 * cv, 0.5 * cv)
        self.grid = ops.BiasGrid([lo], [hi], [400], [0.05 * abs(cv)], W=1e-3 * abs(cv), T_shift=7.0, T=1.0,
                                 stride=w.get("stride", 100), well_tempered=True)

    def step(self):
        cv = self.slab.compute_cv(self.d_pt, self.N_global, self.box)
        bias = self.grid.step(self.t, cv)
        self.slab.forces(self.d_pt, self.N_global, self.box, bias, out=self.d_force)
        self.t += 1

    def algorithmic_bytes(self):
        M = int(np.prod(self.w["mesh"]))
        return 48 * self.N_global + 48 * M


class GraphedStep:
    """CUDA-graph replay of a step whose launch sequence depends only on "deposit step or not": both variants are captured
    once (torch.cuda.graph; the library launches on torch's current stream, the peer-memory / NCCL all-reduce inside is
    capturable) and replayed.  The small workloads are bound by launch latency, not by bandwidth (C2: 3 launches of ~5 us)."""
    use_graph = True

    def _init_graphs(self, stride):
        self._stride, self._graphs, self._eager_done = stride, {}, {True: 0, False: 0}

    def step(self):
        dep = self.t % self._stride == 0
        g = self._graphs.get(dep)
        if g is not None:
            g.replay()
        elif not self.use_graph or getattr(self, "comm_mode", None) == "nccl" or self._eager_done[dep] < 2:
            # (a library collective is left out of graph capture: replaying a captured NCCL all-reduce on 8 ranks hung in this
            #  environment, profiles/r02_notes.md; the peer-memory all-reduce is a plain kernel and replays)
            self._body(self.t)                      # eager: lazy allocations, kernel attributes
            self._eager_done[dep] += 1
        else:
            torch = self.torch
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._body(0 if dep else 1)         # only "deposit or not" reaches the kernels (metad_grid_step)
            self._graphs[dep] = g
            g.replay()
        self.t += 1


def make_comm(comm, mode):
    """Small all-reduces over peer memory (default) or through the library (--comm nccl)."""
    if comm is None or mode == "nccl":
        return comm, "nccl" if comm is not None else None
    from metadynamics_plugin_b200 import sharded
    try:
        return sharded.PeerComm(comm), "p2p"
    except RuntimeError as e:
        if comm.rank == 0:
            print("bench: %s -- falling back to --comm nccl" % e, file=sys.stderr)
        return comm, "nccl"


class LamellarStep(GraphedStep):
    launches_per_step = 3

    def __init__(self, w, ops, torch, comm=None, comm_mode="p2p"):
        self.ops, self.torch, self.w = ops, torch, w
        self.N_global = w["postype"].shape[0]
        self.box = ops.Box.make(w["L"])
        self.comm, self.comm_mode = make_comm(comm, comm_mode)
        pt = w["postype"]
        if comm is not None:            # any particle partition works: contiguous blocks
            pt = np.ascontiguousarray(np.array_split(pt, comm.size)[comm.rank])
            self.launches_per_step = 5 if self.comm_mode == "p2p" else 4       # + finalize, + the all-reduce kernel when it is ours
        self.h_local = pt
        self.N = pt.shape[0]
        self.lam = ops.Lamellar(w["mode"], w["lattice_vectors"])
        self.d_pt = torch.from_numpy(pt).cuda()
        self.d_force = torch.empty_like(self.d_pt)
        g = w["grid"]
        self.ncv = len(g["num_points"])
        self.grid = ops.BiasGrid(g["cv_min"], g["cv_max"], g["num_points"], g["sigma"], W=w["W"], T_shift=w["deltaT"], T=w["T"],
                                 stride=w["stride"], well_tempered=True)
        self.cvs = torch.zeros(self.ncv, dtype=torch.float64, device="cuda")
        if self.ncv == 2:       # second CV = aspect ratio Lx/Ly of the (cubic) box: host scalar
            self.cvs[1] = 1.0
        self.t = 0
        self._init_graphs(w["stride"])

    def sharded_cv(self):
        # partial modes -> all-reduce of 2*n_wave doubles -> CV (LamellarOrderParameterGPU.cc:70-77)
        self.lam.compute_modes(self.d_pt, self.N_global, self.box, finalize=False)
        self.comm.all_reduce_sum(self.lam.modes)
        return self.lam.finalize(self.N_global)

    def _body(self, t):
        if self.comm is None:
            cv = self.lam.compute_modes(self.d_pt, self.N_global, self.box)
        else:
            cv = self.sharded_cv()
        if self.ncv == 1:
            bias = self.grid.step(t, cv)
        else:
            self.cvs[0:1].copy_(cv)
            bias = self.grid.step(t, self.cvs)
        self.lam.forces(self.d_pt, self.N_global, self.box, bias[0:1], out=self.d_force)

    def algorithmic_bytes(self):
        return 48 * self.N_global


class WteStep(GraphedStep):
    """WellTemperedEnsemble (cv.potential_energy) over particle shards: reduce the potential energy (+ all-reduce of one
    double, WellTemperedEnsemble.cc:58-64), 1-D grid bias, scale net force / torque / virial by 1 + bias (:135-188)."""
    launches_per_step = 3

    def __init__(self, w, ops, torch, comm=None, comm_mode="p2p"):
        self.ops, self.torch, self.w = ops, torch, w
        self.N_global = w["net_force"].shape[0]
        self.comm, self.comm_mode = make_comm(comm, comm_mode)
        part = slice(None) if comm is None else np.array_split(np.arange(self.N_global), comm.size)[comm.rank]
        if comm is not None:
            self.launches_per_step = 4 if self.comm_mode == "p2p" else 3
        self.h_local = np.ascontiguousarray(w["net_force"][part])
        self.N = self.h_local.shape[0]
        self.pitch = (self.N + 15) // 16 * 16
        self.d_pt = torch.from_numpy(self.h_local).cuda()            # "d_pt" = the array the e2e leg uploads: the net force
        self.d_force = self.d_pt                                       # scaled in place: the array the e2e leg reads back
        self.d_tq = torch.zeros_like(self.d_pt)
        self.d_vir = torch.zeros(6 * self.pitch, dtype=torch.float32, device="cuda")
        self.pe = torch.zeros(1, dtype=torch.float64, device="cuda")
        pe0 = float(w["net_force"][:, 3].astype(np.float64).sum())
        self.grid = ops.BiasGrid([pe0 - 0.2 * abs(pe0)], [pe0 + 0.2 * abs(pe0)], [400], [0.01 * abs(pe0)], W=1e-6, T_shift=7.0, T=1.0,
                                 stride=w["stride"], well_tempered=True)
        self.t = 0
        self._init_graphs(w["stride"])

    def _body(self, t):
        self.ops.wte_reduce(self.d_pt, 0.0, out=self.pe)
        if self.comm is not None:
            self.comm.all_reduce_sum(self.pe)
        bias = self.grid.step(t, self.pe)
        self.ops.wte_scale(self.d_pt, self.d_tq, self.d_vir, self.pitch, bias)

    def algorithmic_bytes(self):
        return (16 + 2 * (16 + 16 + 24)) * self.N_global


def run_ours(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    from metadynamics_plugin_b200 import ops

    if world > 1:
        import torch.distributed as dist
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=120))
    peak, peak_src = load_peaks()
    w = make_workload(args.workload)
    comm = None
    if world > 1:
        from metadynamics_plugin_b200 import sharded
        comm = sharded.TorchComm()
    GraphedStep.use_graph = not args.no_graph
    if w["kind"] == "mesh":
        runner = MeshStep(w, ops, torch) if world == 1 else MeshSlabStep(w, ops, torch, comm, mode=args.comm, sync=args.p2p_sync)
    elif w["kind"] == "wte":
        runner = WteStep(w, ops, torch, comm, args.comm)
    else:
        runner = LamellarStep(w, ops, torch, comm, args.comm)

    for kv in args.late_knob:               # experiments that break the results (timing only): set after the calibration
        k, v = (int(x) for x in kv.split("="))
        runner.mesh.set(k, v)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()          # sampled from warm-up to the end of the e2e leg (all under load); see clocks.window
    t_load0 = time.time()
    rounds = 0
    while True:                  # W warm-up steps, and at least ~0.3 s of load so the clock sampler has data
        for _ in range(args.warmup):
            runner.step()
        torch.cuda.synchronize()
        rounds += 1
        # multi-rank: every rank must issue the same number of collectives -> a fixed number of rounds, no local clock
        if (world == 1 and time.time() - t_load0 > 0.3) or (world > 1 and rounds >= 8):
            break
    sync_all()
    rebuilds_before = runner.mesh.stats()["rebuilds"] if w["kind"] == "mesh" else 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t_host0 = time.perf_counter()
    for _ in range(args.steps):
        runner.step()
    host_enqueue_ms = (time.perf_counter() - t_host0) * 1e3 / args.steps      # CPU time to enqueue one step (no sync inside)
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1)
    rebuilds_timed = 0
    if w["kind"] == "mesh":
        rebuilds_timed = runner.mesh.stats()["rebuilds"] - rebuilds_before
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    ms_per_step = ms / args.steps

    # per-kernel timing of the dominant kernel (CUDA events inside the library, on the launching stream)
    roofline = None
    if w["kind"] == "mesh" and world > 1:
        achieved = runner.algorithmic_bytes() / (ms_per_step * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": "whole sharded step (%s)" % ("compute stages with the transposes fused in, halo pushes, flag barriers over peer memory"
                                                                          if runner.comm_mode == "p2p" else "compute stages + NCCL collectives"), "achieved": achieved,
                    "peak": peak * world, "unit": "GB/s", "frac": achieved / (peak * world), "traffic": None,
                    "peak_source": peak_src + " x %d GPUs" % world}
        if runner.comm_mode == "p2p" and args.p2p_sync == "barrier":
            # where the sharded step spends its time: events around the 12 segments of the peer-memory step (no graph
            # replay while profiling); a barrier's time is the wait for the slowest rank plus the flag round trip
            runner.mesh.set(2, 1)
            acc = {}
            nprof = 10
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            for _ in range(nprof):
                # the step of MeshSlabStep.step, with events around the two parts outside the library's own segment timers:
                # the (replicated) grid step and the gather, so that the segments add up to the step
                cv = runner.slab.compute_cv(runner.d_pt, runner.N_global, runner.box)
                ev[0].record(); bias = runner.grid.step(runner.t, cv); ev[1].record()
                ev[2].record(); runner.slab.forces(runner.d_pt, runner.N_global, runner.box, bias, out=runner.d_force); ev[3].record()
                runner.t += 1
                torch.cuda.synchronize()
                for k, v in runner.mesh.p2p_timings().items():
                    acc[k] = acc.get(k, 0.0) + v / nprof
                acc["grid_step"] = acc.get("grid_step", 0.0) + ev[0].elapsed_time(ev[1]) / nprof
                acc["gather"] = acc.get("gather", 0.0) + ev[2].elapsed_time(ev[3]) / nprof
            acc["sum_of_segments"] = sum(acc.values())     # eager launches with events in between: larger than the graph-replayed step
            runner.mesh.set(2, 0)
            keys = list(acc)
            t = torch.tensor([acc[k] for k in keys], dtype=torch.float64, device="cuda")
            tmax = t.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            roofline["segment_ms_rank0"] = {k: round(v, 4) for k, v in zip(keys, t.tolist())}
            roofline["segment_ms_max_over_ranks"] = {k: round(v, 4) for k, v in zip(keys, tmax.tolist())}
    elif w["kind"] == "mesh":
        runner.mesh.set(2, 1)
        acc = {}
        nprof = min(args.steps, 10)
        for _ in range(nprof):
            runner.step()
            torch.cuda.synchronize()
            for k, v in runner.mesh.timings().items():
                acc[k] = acc.get(k, 0.0) + v / nprof
        runner.mesh.set(2, 0)
        sb = runner.stage_bytes()
        top = max((k for k in acc if sb[k] > 0), key=lambda k: acc[k])
        achieved = sb[top] / (acc[top] * 1e-3) / 1e9
        traffic, traffic_src = None, None
        try:        # DRAM bytes per launch of that kernel from the committed ncu --set full capture of this workload
            tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            if tj["workload"] == w["name"]:
                traffic, traffic_src = tj["dram_bytes_per_launch"].get(top), tj["source"]
        except Exception:
            pass
        roofline = {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes": sb[top], "peak_source": peak_src,
                    "stage_ms": {k: round(v, 4) for k, v in acc.items()},
                    "step_frac": runner.algorithmic_bytes() / (ms_per_step * 1e-3) / 1e9 / peak}
    else:
        achieved = runner.algorithmic_bytes() / (ms_per_step * 1e-3) / 1e9
        kern = "wte_reduce+grid_step+wte_scale (whole step)" if w["kind"] == "wte" else "lamellar_modes+grid_step+lamellar_force (whole step)"
        if world > 1:
            kern += " + one all-reduce of %s doubles (%s)" % ("1" if w["kind"] == "wte" else "2 n_q", "peer memory, csrc/peer.cu" if runner.comm_mode == "p2p" else "NCCL")
        roofline = {"bound": "hbm", "kernel": kern, "achieved": achieved, "peak": peak * world,
                    "unit": "GB/s", "frac": achieved / (peak * world), "traffic": None,
                    "peak_source": peak_src + (" x %d GPUs" % world if world > 1 else ""),
                    "cuda_graph": bool(GraphedStep.use_graph)}

    extra = {}
    if w["kind"] == "mesh" and world == 1 and not args.no_extra_legs:
        extra = mesh_extra_legs(args, w, runner, ops, torch, ms_per_step, rebuilds_timed)
        ms_per_step = extra["tile_order"]["ms_per_step_steady_state"]
        roofline["step_frac"] = runner.algorithmic_bytes() / (ms_per_step * 1e-3) / 1e9 / peak

    # end to end through the same calls with host buffers
    h_pt = torch.from_numpy(getattr(runner, "h_local", w["postype"])).pin_memory()
    h_force = torch.empty_like(h_pt).pin_memory()
    h_cv = torch.zeros(1, dtype=torch.float64).pin_memory()
    cv_t = runner.mesh.cv if w["kind"] == "mesh" else (runner.pe if w["kind"] == "wte" else runner.lam.cv)
    if world > 1:
        h2d = torch.tensor([float(h_pt.numel() * 4), float(h_force.numel() * 4 + 8)], dtype=torch.float64, device="cuda")
        dist.all_reduce(h2d)
        h2d_bytes, d2h_bytes = int(h2d[0].item()), int(h2d[1].item())
    else:
        h2d_bytes, d2h_bytes = int(h_pt.numel() * 4), int(h_force.numel() * 4 + 8)
    e2e_steps = max(1, min(args.steps, 20))

    def e2e_step():
        runner.d_pt.copy_(h_pt, non_blocking=True)
        runner.step()
        h_force.copy_(runner.d_force, non_blocking=True)
        h_cv.copy_(cv_t, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    e2e_step()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    sync_all()
    e2e_serial_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps

    # the same work as a pipeline: the upload of step i+1 and the read-back of step i-1 overlap step i (two copy streams, staging
    # buffers on the device, two pinned result buffers); every step still uploads its inputs and reads back forces + CV
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    stage_in = [torch.empty_like(runner.d_pt) for _ in range(2)]
    stage_out = [torch.empty_like(runner.d_force) for _ in range(2)]
    stage_cv = [torch.zeros(1, dtype=torch.float64, device="cuda") for _ in range(2)]
    h_forces = [h_force, torch.empty_like(h_force).pin_memory()]
    h_cvs = [h_cv, torch.zeros(1, dtype=torch.float64).pin_memory()]
    ev = {k: [torch.cuda.Event() for _ in range(2)] for k in ("in", "free", "done", "out")}

    def e2e_pipeline(K):
        main = torch.cuda.current_stream()
        for b in range(2):
            ev["free"][b].record(main); ev["out"][b].record(main)
        with torch.cuda.stream(s_in):
            stage_in[0].copy_(h_pt, non_blocking=True); ev["in"][0].record(s_in)
        for i in range(K):
            b = i & 1
            if i + 1 < K:
                with torch.cuda.stream(s_in):
                    s_in.wait_event(ev["free"][1 - b])
                    stage_in[1 - b].copy_(h_pt, non_blocking=True); ev["in"][1 - b].record(s_in)
            main.wait_event(ev["in"][b])
            runner.d_pt.copy_(stage_in[b]); ev["free"][b].record(main)
            runner.step()
            main.wait_event(ev["out"][b])
            stage_out[b].copy_(runner.d_force); stage_cv[b].copy_(cv_t); ev["done"][b].record(main)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev["done"][b])
                h_forces[b].copy_(stage_out[b], non_blocking=True); h_cvs[b].copy_(stage_cv[b], non_blocking=True); ev["out"][b].record(s_out)
        torch.cuda.synchronize()

    e2e_pipeline(2)
    sync_all()
    t0 = time.perf_counter()
    e2e_pipeline(e2e_steps)
    sync_all()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    if world > 1:
        t = torch.tensor([e2e_ms, e2e_serial_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms, e2e_serial_ms = t.tolist()
    del stage_in, stage_out

    n_global = w["postype"].shape[0]
    clocks = None
    if rank == 0:
        clocks = sampler.stop()
        clocks["window"] = "warm-up + timed steps + per-kernel timing + e2e leg (continuous load)"
    out = {
        "metric": "cv_bias_force_steps_per_sec", "value": 1e3 / ms_per_step, "unit": "steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32 (fixed-point density, fp64 accumulation of the CV)", "data": "synthetic",
        "ns_per_particle_step": ms_per_step * 1e6 / n_global,
        "config": config_for(w, world),
        "run_info": {"parallelism": "1 GPU" if world == 1 else (("z-slab mesh + particle sharding over %d GPUs, %s" % (
                       world, "peer memory over NVLink: transposes fused into the FFT sweeps, pushed halos, flag barriers (no NCCL call in a step; "
                       "CV agrees with the NCCL path to %.1e)" % runner.cv_check if runner.comm_mode == "p2p" else "NCCL all-to-all / halo exchange / all-reduce"))
                                                              if w["kind"] == "mesh" else "particles sharded over %d GPUs, one all-reduce of a few doubles per step (%s)" % (
                                                                  world, "peer memory over NVLink, csrc/peer.cu; step replayed from a CUDA graph" if runner.comm_mode == "p2p" else "NCCL")),
                     "tile_order_rebuild_period": getattr(runner, "period", None), "tile_order_rebuilds_in_timed_region": rebuilds_timed},
        "roofline": roofline,
        "e2e": {"value": 1e3 / e2e_ms, "unit": "steps/s", "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": d2h_bytes, "steps": e2e_steps, "serial_value": 1e3 / e2e_serial_ms,
                "note": "value: upload of step i+1 and read-back of step i-1 overlap step i (every step uploads its inputs from pinned "
                        "memory and reads back forces + CV); serial_value: copy in, step, copy out, synchronise"},
        "host_enqueue_ms_per_step": host_enqueue_ms,
        "gpu_launches": runner.launches_per_step * args.steps + getattr(runner, "launches_per_rebuild", 0) * rebuilds_timed,
        "clocks": clocks,
    }
    out.update(extra)
    # the same fractions against the nominal 8 TB/s of the north star (SURVEY 8d names both denominators)
    roofline["frac_nominal_8TBps"] = roofline["achieved"] / (8000.0 * world)
    if "step_frac" in roofline:
        roofline["step_frac_nominal_8TBps"] = roofline["step_frac"] * peak / 8000.0
    parity = None if args.no_parity else parity_block(w, runner, world, rank, torch)
    if world > 1:
        dist.barrier()
    if rank == 0:
        if parity is not None:
            out["parity"] = parity
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(w, budget_s=args.cpu_budget)
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    if rank == 0 and parity is not None and not parity["ok"]:
        print("bench: PARITY FAILURE against the oracle: %r" % parity, file=sys.stderr)
        sys.exit(3)


def config_for(w, world):
    """The `config` object of the JSON line: identical for both arms (ours / --impl reference) of the same workload."""
    return {"workload": "%s: %s" % (w["name"], describe(w)) + (", " + w["order"] if "order" in w else ""), "N": int(w["postype"].shape[0]),
            "l2": "inputs larger than L2 (no flush)" if w["postype"].nbytes > 126e6 else "inputs smaller than L2 (small workload, latency-bound; no flush)",
            "gpus": world}


def describe(w):
    if w["kind"] == "wte":
        return "WellTemperedEnsemble (potential energy) + 1-D well-tempered grid bias, N=%d" % w["net_force"].shape[0]
    if w["kind"] == "mesh":
        return "OrderParameterMesh CV + 1-D well-tempered grid bias, N=%d, mesh %dx%dx%d, L=%.3f" % (
            w["postype"].shape[0], *w["mesh"], w["L"])
    return "LamellarOrderParameter %d wave vectors + %d-D well-tempered grid bias, N=%d, L=%.3f" % (
        len(w["lattice_vectors"]), len(w["grid"]["num_points"]), w["postype"].shape[0], w["L"])


# ---------------------------------------------------------------------------------------------------- extra legs
def _time_steps(torch, fn, n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def mesh_extra_legs(args, w, runner, ops, torch, ms_measured, rebuilds_timed):
    """What the headline number leaves out, measured in the same run:
      tile_order   cost of one rebuild of the tile order and the steady-state step with it amortised over the period;
      orders       the same step for other particle orders of the input (random; Morton order on 4^3-cell blocks);
      moving       particles displaced by a Gaussian of 0.05 cell per step (stale order, drift counters, rebuilds);
      host_class   the step through the reference-facing classes (cv.mesh + integrate.mode_metadynamics -> _metadynamics)."""
    out = {}
    period = runner.period
    # -- rebuild cost: every call rebuilds vs no call rebuilds
    runner.mesh.set(0, 1)
    runner.step()
    t_every = _time_steps(torch, runner.step, 6)
    runner.mesh.set(0, 1 << 30)
    runner.step()
    t_never = _time_steps(torch, runner.step, 12)
    runner.mesh.set(0, period)
    rebuild_ms = max(t_every - t_never, 0.0)
    steady = (ms_measured * args.steps - rebuilds_timed * rebuild_ms) / args.steps + rebuild_ms / period
    out["tile_order"] = {"rebuild_ms": round(rebuild_ms, 4), "period": period, "amortised_ms_per_step": round(rebuild_ms / period, 5),
                         "rebuilds_in_timed_region": rebuilds_timed, "ms_per_step_timed_region": ms_measured,
                         "ms_per_step_steady_state": steady,
                         "note": "ms_per_step / value = timed region with its rebuilds replaced by rebuild_ms / period per step"}
    # -- other particle orders of the same configuration
    orders = {"cell_sorted": round(t_never + rebuild_ms / period, 4)}
    for order in ("sfc", "random"):
        w2 = make_workload(args.workload, order=order)
        r2 = MeshStep(w2, ops, torch, period=period)
        for _ in range(3):
            r2.step()
        t = _time_steps(torch, r2.step, 2 * period)             # two periods: contains two rebuilds
        r2.mesh.set(2, 1)
        r2.step(); torch.cuda.synchronize()
        st = r2.mesh.timings()
        r2.mesh.set(2, 0)
        orders[order] = round(t, 4)
        orders[order + "_stage_ms"] = {k: round(v, 4) for k, v in st.items()}
        del r2, w2
        torch.cuda.empty_cache()
    out["orders_ms_per_step"] = orders
    # -- moving particles: Gaussian displacement of 0.05 cell per step, periodic wrap, 2 periods
    mv = MeshStep(w, ops, torch, period=period)
    L, h = float(w["L"]), float(w["L"]) / w["mesh"][0]
    half = torch.tensor(L / 2, dtype=torch.float32, device="cuda")
    for _ in range(3):
        mv.step()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(2 * period)]
    drift_max, rb0 = 0, mv.mesh.stats()["rebuilds"]
    gen = torch.Generator(device="cuda"); gen.manual_seed(1)
    for a, b in ev:
        xyz = mv.d_pt[:, :3]
        xyz.add_(torch.randn(xyz.shape, device="cuda", generator=gen) * (0.05 * h))
        xyz.copy_(torch.remainder(xyz + half, 2 * half) - half)
        xyz.clamp_(min=-L / 2, max=float(np.nextafter(np.float32(L / 2), np.float32(0))))
        a.record(); mv.step(); b.record()
    torch.cuda.synchronize()
    st = mv.mesh.stats()
    out["moving"] = {"displacement_cells_per_step": 0.05, "steps": len(ev), "ms_per_step": round(sum(a.elapsed_time(b) for a, b in ev) / len(ev), 4),
                     "rebuilds": st["rebuilds"] - rb0, "drifted_particles_last_step": st["drifted"],
                     "note": "the tile order goes stale between rebuilds: drifted particles take the direct path (global atomics / loads)"}
    del mv
    torch.cuda.empty_cache()
    # -- the reference-facing host classes
    try:
        out["host_class"] = host_class_leg(w, torch, min(args.steps, 20))
    except Exception as e:       # a failure here must not take the headline down with it, but it must be visible
        out["host_class"] = {"error": repr(e)}
    return out


def host_class_leg(w, torch, steps):
    """cv.mesh + integrate.mode_metadynamics (the reference's Python API over the pybind `_metadynamics` classes):
    device-resident steps of IntegratorMetaDynamics::update, and the same with host arrays in and out per step."""
    from metadynamics_plugin_b200 import cv, integrate, hoomd_shim as hoomd
    hoomd.context.initialize()
    pt = w["postype"]
    N = pt.shape[0]
    sd = hoomd.init.from_arrays(pt[:, :3], pt[:, 3].view(np.int32), ["A"], w["L"])
    meta = integrate.mode_metadynamics(dt=0.005, mode="well_tempered", stride=w.get("stride", 100), deltaT=7.0, W=1e-12)
    mesh = cv.mesh(nx=w["mesh"][0], ny=w["mesh"][1], nz=w["mesh"][2], mode={"A": float(w["mode"][0])})
    val = mesh.cpp_force.getCurrentValue(0)
    mesh.set_grid(cv_min=0.5 * val, cv_max=1.5 * val, num_points=400)
    mesh.sigma = 0.05 * abs(val)
    hoomd.run(3)                                                     # prepRun + warm-up
    integ = hoomd.context.current.integrator.cpp_integrator
    cur = hoomd.context.current
    t0 = time.perf_counter()
    for _ in range(steps):
        integ.update(cur.timestep); cur.timestep += 1
    torch.cuda.synchronize()
    dev_ms = (time.perf_counter() - t0) * 1e3 / steps
    pd = sd.getParticleData()
    t0 = time.perf_counter()
    for _ in range(steps):
        pd.setPositions(pt)                                          # host -> device (pageable numpy memory: what the API takes)
        integ.update(cur.timestep); cur.timestep += 1
        f = mesh.get_forces()                                        # device -> host
        cvv = mesh.cpp_force.getLogValue("cv_mesh", cur.timestep)
    e2e_ms = (time.perf_counter() - t0) * 1e3 / steps
    return {"api": "cv.mesh + integrate.mode_metadynamics -> _metadynamics.OrderParameterMeshGPU / IntegratorMetaDynamics.update",
            "ms_per_step_device_resident": round(dev_ms, 4), "e2e_ms_per_step_host_arrays": round(e2e_ms, 3), "steps": steps,
            "cv": float(cvv), "force_absmax": float(np.abs(f).max())}


# ---------------------------------------------------------------------------------------------------- parity
CV_TOL, FORCE_TOL = 1e-6, 1e-5      # north star: CV 1e-6 relative, forces 1e-5 of max|F|, cell indices bit-exact


def parity_block(w, runner, world, rank, torch):
    """The bench's own workload against the double oracle, in the same run (the checker, never the thing measured):
    one extra device evaluation with bias = 1 next to one oracle step.  Multi-GPU: every rank evaluates its shard, rank 0
    compares the (global) CV and the forces / cells of its own shard.  Raises SystemExit(3) outside the tolerances."""
    one = torch.tensor([1.0], dtype=torch.float64, device="cuda")
    pt_local = getattr(runner, "h_local", w["postype"])
    n_global = w["postype"].shape[0]
    out = {"oracle": "f64 instance of the CPU restatement (pinned to the reference's own sources, tests/test_reference_build.py); "
                     "mesh forces with |x| exact instead of the double build's copysignf rounding (DESIGN.md section 2)",
           "cv_tol": CV_TOL, "force_tol": FORCE_TOL, "shard": "rank 0 of %d" % world if world > 1 else "all particles"}
    if w["kind"] == "mesh":
        m = runner.mesh
        if world == 1:
            m.set(3, 1)                                  # keep the cell indices of this evaluation
            cv = m.compute_cv(runner.d_pt, runner.N, runner.box).cpu().item()
            f = m.forces(runner.d_pt, runner.N, runner.box, one).cpu().numpy()
            cells = m.cells()
            m.set(3, 0)
        else:
            m.set(3, 1)
            cv = runner.slab.compute_cv(runner.d_pt, n_global, runner.box).cpu().item()
            f = runner.slab.forces(runner.d_pt, n_global, runner.box, one).cpu().numpy()
            cells = m.cells()
            m.set(3, 0)
        if rank != 0:
            return None
        from oracle import pyoracle as po
        t0 = time.perf_counter()
        o32 = po.Mesh(*w["mesh"], w["mode"], w["L"], n_global, "f32")
        o32.assign(pt_local)
        out["cells_bitexact"] = bool(np.array_equal(cells, o32.cells()))
        del o32
        o = po.Mesh(*w["mesh"], w["mode"], w["L"], n_global, "f64", literal_copysignf=False)
        cvo = o.current_value(w["postype"])
        fo = o.forces(pt_local, 1.0)
        out["oracle_seconds"] = round(time.perf_counter() - t0, 2)
    elif w["kind"] == "wte":
        nf = torch.from_numpy(pt_local).cuda()
        pe = runner.ops.wte_reduce(nf, 0.0)
        if runner.comm is not None:
            runner.comm.all_reduce_sum(pe)
        cv = pe.cpu().item()
        half = torch.tensor([0.5], dtype=torch.float64, device="cuda")
        runner.ops.wte_scale(nf, runner.d_tq, runner.d_vir, runner.pitch, half)
        f = nf.cpu().numpy()
        if rank != 0:
            return None
        from oracle import pyoracle as po
        t0 = time.perf_counter()
        cvo = po.wte_pe(w["net_force"], 0.0)
        fo, _, _, _ = po.wte_scale(pt_local, np.zeros_like(pt_local), np.zeros(6 * runner.pitch, np.float32), runner.pitch, 0.5, np.zeros(6))
        out["cells_bitexact"] = None
        out["oracle_seconds"] = round(time.perf_counter() - t0, 2)
    else:
        lam = runner.lam
        if runner.comm is None:
            cv = lam.compute_modes(runner.d_pt, n_global, runner.box).cpu().item()
        else:
            cv = runner.sharded_cv().cpu().item()
        f = lam.forces(runner.d_pt, n_global, runner.box, one).cpu().numpy()
        if rank != 0:
            return None
        from oracle import pyoracle as po
        t0 = time.perf_counter()
        cvo, _ = po.lamellar_cv(w["postype"], n_global, w["mode"], w["lattice_vectors"], w["L"])
        fo = po.lamellar_forces(pt_local, n_global, w["mode"], w["lattice_vectors"], w["L"], 1.0)
        out["cells_bitexact"] = None                     # no integer work on the Lamellar path
        out["oracle_seconds"] = round(time.perf_counter() - t0, 2)
    out["cv"] = cv
    out["cv_oracle"] = cvo
    out["cv_rel"] = abs(cv - cvo) / abs(cvo) if cvo != 0 else abs(cv)
    fmax = float(np.abs(fo).max())
    out["force_rel_max"] = float(np.abs(f - fo).max() / fmax) if fmax > 0 else float(np.abs(f).max())
    out["ok"] = bool(out["cv_rel"] <= CV_TOL and out["force_rel_max"] <= FORCE_TOL and out["cells_bitexact"] in (True, None))
    return out


# ---------------------------------------------------------------------------------------------------- CPU side
def reference_step_fn(w, prec="f32"):
    """One full CV + bias-force step through the REFERENCE'S OWN classes (oracle/_ref: OrderParameterMesh.cc /
    LamellarOrderParameter.cc compiled unmodified from the reference's sources against the HOOMD stand-in, prebuilt by
    __graft_entry__.build() where the reference is mounted): getCurrentValue + setBiasFactor + computeBiasForces on a resident
    particle set; the (tiny) grid step of the Lamellar workloads comes from the oracle, which equals the reference's integrator
    bit for bit (tests/test_reference_build.py).  None if oracle/_ref is not there or the workload has no such class."""
    if w["kind"] not in ("mesh", "lamellar"):
        return None
    try:
        from oracle import pyref, pyoracle as po
        if not pyref.available():
            return None
        phases = {}
        if w["kind"] == "mesh":
            plan = pyref.StepPlan("mesh", w["postype"], w["L"], w["mode"], prec, dims=w["mesh"])

            def step():
                cv, t_cv, t_f = plan.step(1.0)
                phases["getCurrentValue"] = phases.get("getCurrentValue", 0.0) + t_cv
                phases["computeBiasForces"] = phases.get("computeBiasForces", 0.0) + t_f
                return cv
            return step, phases
        plan = pyref.StepPlan("lamellar", w["postype"], w["L"], w["mode"], prec, lattice_vectors=w["lattice_vectors"])
        g = w["grid"]
        grid = po.Grid(g["cv_min"], g["cv_max"], g["num_points"], g["sigma"], W=w["W"], T_shift=w["deltaT"], T=w["T"], stride=w["stride"],
                       well_tempered=True, prec=prec)
        state = {"t": 0, "bias": 1.0}

        def step():
            # the bias factor of the previous grid step scales the forces (same arithmetic, one call into the reference's class)
            cv, t_cv, t_f = plan.step(state["bias"])
            t0 = time.perf_counter()
            vals = [cv] + ([1.0] if len(g["num_points"]) == 2 else [])
            state["bias"] = float(grid.update(state["t"], vals)[0])
            state["t"] += 1
            for k, v in (("getCurrentValue", t_cv), ("grid (oracle)", time.perf_counter() - t0), ("computeBiasForces", t_f)):
                phases[k] = phases.get(k, 0.0) + v
            return cv
        return step, phases
    except Exception as e:          # a missing or unloadable oracle/_ref must not take the arm down: the port is always there
        print("reference build not usable (%s): falling back to the oracle port" % e, file=sys.stderr)
        return None


def cpu_arm(w, prec="f32"):
    """(step, phases, kind, description) of the CPU arm: the reference's own code if oracle/_ref is built, else the oracle port."""
    ref = reference_step_fn(w, prec)
    if ref is not None:
        return ref[0], ref[1], "reference", ("the reference's own CPU classes (oracle/_ref: its .cc files compiled unmodified, "
                                             "-O3 -march=x86-64-v3, float build), 1 thread (the reference CPU path is serial per MPI rank)")
    step, phases = cpu_step_fn(w, prec)
    return step, phases, "port", "oracle port of the reference CPU path, float instance, 1 thread (reference CPU path is serial per MPI rank)"


def cpu_step_fn(w, prec="f32"):
    """One full CV + bias-force step of the reference's CPU path (oracle port, single-threaded like the reference)."""
    from oracle import pyoracle as po
    N = w["postype"].shape[0]
    if w["kind"] == "mesh":
        m = po.Mesh(*w["mesh"], w["mode"], w["L"], N, prec)
        phases = {}

        def step():
            t0 = time.perf_counter(); m.assign(w["postype"])
            t1 = time.perf_counter(); m.update()
            t2 = time.perf_counter(); cv = m.cv()
            t3 = time.perf_counter(); m.forces(w["postype"], 1.0)
            t4 = time.perf_counter()
            for k, v in (("assign", t1 - t0), ("fft+convolve", t2 - t1), ("cv_sum", t3 - t2), ("forces", t4 - t3)):
                phases[k] = phases.get(k, 0.0) + v
            return cv
        return step, phases
    if w["kind"] == "wte":
        nf = w["net_force"]
        pe0 = float(nf[:, 3].astype(np.float64).sum())
        grid = po.Grid([pe0 - 0.2 * abs(pe0)], [pe0 + 0.2 * abs(pe0)], [400], [0.01 * abs(pe0)], W=1e-6, T_shift=7.0, T=1.0, stride=w["stride"],
                       well_tempered=True, prec=prec)
        pitch = (N + 15) // 16 * 16
        tq, vir = np.zeros_like(nf), np.zeros(6 * pitch, np.float32)
        state, phases = {"t": 0}, {}

        def step():
            t0 = time.perf_counter()
            pe = po.wte_pe(nf, 0.0, prec)
            t1 = time.perf_counter()
            b = grid.update(state["t"], [pe])
            t2 = time.perf_counter()
            po.wte_scale(nf, tq, vir, pitch, b[0], np.zeros(6), prec)
            t3 = time.perf_counter()
            state["t"] += 1
            for k, v in (("reduce", t1 - t0), ("grid", t2 - t1), ("scale", t3 - t2)):
                phases[k] = phases.get(k, 0.0) + v
            return pe
        return step, phases
    g = w["grid"]
    grid = po.Grid(g["cv_min"], g["cv_max"], g["num_points"], g["sigma"], W=w["W"], T_shift=w["deltaT"], T=w["T"], stride=w["stride"],
                   well_tempered=True, prec=prec)
    state = {"t": 0}
    phases = {}

    def step():
        t0 = time.perf_counter()
        cv, _ = po.lamellar_cv(w["postype"], N, w["mode"], w["lattice_vectors"], w["L"], prec)
        t1 = time.perf_counter()
        vals = [cv] + ([1.0] if len(g["num_points"]) == 2 else [])
        b = grid.update(state["t"], vals)
        t2 = time.perf_counter()
        po.lamellar_forces(w["postype"], N, w["mode"], w["lattice_vectors"], w["L"], b[0], prec)
        t3 = time.perf_counter()
        state["t"] += 1
        for k, v in (("cv", t1 - t0), ("grid", t2 - t1), ("forces", t3 - t2)):
            phases[k] = phases.get(k, 0.0) + v
        return cv
    return step, phases


def cpu_baseline(w, budget_s=25.0):
    step, phases, kind, what = cpu_arm(w)
    n, t0 = 0, time.perf_counter()
    while True:
        step()
        n += 1
        if time.perf_counter() - t0 > budget_s or n >= 50:
            break
    dt = (time.perf_counter() - t0) / n
    return {"value": 1.0 / dt, "unit": "steps/s", "cores": 1, "kind": kind,
            "sample": "%d full step(s) of the same workload; %s; host has %d cores" % (n, what, os.cpu_count()),
            "phases_s_per_step": {k: v / n for k, v in phases.items()}}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path -- its own classes from oracle/_ref where that is
    built (the full plugin needs an installed HOOMD-blue and cannot be built here), else the oracle port."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = make_workload(args.workload)
    step, phases, kind, what = cpu_arm(w)
    budget = 200.0                       # seconds for the whole arm: a full C4 step takes ~40 s in the reference's own code, 7.5 s in the port
    t_start = time.perf_counter()
    done_w = 0
    for _ in range(args.warmup):
        if done_w >= 1 and time.perf_counter() - t_start > 0.08 * budget:
            break
        step(); done_w += 1
    phases.clear()
    n, t0 = 0, time.perf_counter()
    for _ in range(args.steps):
        step(); n += 1
        if time.perf_counter() - t_start > budget:
            break
    dt = (time.perf_counter() - t0) / n
    N = w["postype"].shape[0]
    sample = ("%d of %d requested full step(s) (time-bounded), %d warm-up; %s; host has %d cores" % (n, args.steps, done_w, what, os.cpu_count()))
    out = {"impl": "reference", "metric": "cv_bias_force_steps_per_sec", "value": 1.0 / dt, "unit": "steps/s",
           "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": n, "warmup": done_w, "ms_per_step": dt * 1e3,
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "ns_per_particle_step": dt * 1e9 / N,
           "config": config_for(w, int(os.environ.get("WORLD_SIZE", "1"))),
           "cpu_baseline": {"value": 1.0 / dt, "unit": "steps/s", "cores": 1, "kind": kind, "sample": sample,
                            "phases_s_per_step": {k: v / n for k, v in phases.items()}},
           "e2e": {"value": 1.0 / dt, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="C4", choices=["C1", "C2", "C3", "C4", "C5", "WTE"])
    ap.add_argument("--no-graph", action="store_true", help="Lamellar / WTE workloads: launch the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--comm", default="p2p", choices=["p2p", "nccl"], help="multi-GPU mesh path: peer memory (default) or NCCL collectives")
    ap.add_argument("--p2p-sync", default="barrier", choices=["fused", "barrier"],
                    help="peer-memory mode: separate barrier launches (default, measured faster) or inter-rank signal/wait inside the kernels")
    ap.add_argument("--tile-order", default=None, choices=["bank", "layer"], help="order of the particles inside a tile (single GPU; default: library default = bank)")
    ap.add_argument("--order", default="sorted", choices=["sorted", "random", "sfc"], help="C3 / C4: particle order of the input (cell-sorted headline, random stress case)")
    ap.add_argument("--merge-push", action="store_true", help="peer-memory mode: halo push and barrier in one launch (measured slower)")
    ap.add_argument("--no-pdl", action="store_true", help="launch the per-step kernels without programmatic dependent launch")
    ap.add_argument("--mesh-knob", action="append", default=[], metavar="K=V", help="metad_mesh_set(plan, K, V) before the first step (experiments)")
    ap.add_argument("--late-knob", action="append", default=[], metavar="K=V", help="like --mesh-knob, applied after the set-up (timing experiments)")
    ap.add_argument("--no-extra-legs", action="store_true", help="skip the rebuild-cost / particle-order / moving-particle / host-class legs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle parity block (experiments only)")
    ap.add_argument("--cpu-budget", type=float, default=25.0)
    args = ap.parse_args()
    global TILE_ORDER, NO_PDL, PARTICLE_ORDER, MERGE_PUSH, MESH_KNOBS
    MESH_KNOBS = [tuple(int(x) for x in kv.split("=")) for kv in args.mesh_knob]
    MERGE_PUSH = args.merge_push
    PARTICLE_ORDER = args.order
    TILE_ORDER = args.tile_order
    NO_PDL = args.no_pdl
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
