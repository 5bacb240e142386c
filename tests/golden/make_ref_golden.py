"""Generates tests/golden/ref_golden.npz from the REFERENCE's own code.

The reference's CPU classes (CollectiveVariable.cc, LamellarOrderParameter.cc, OrderParameterMesh.cc, AspectRatio.cc,
IndexGrid.cc, IntegratorMetaDynamics.cc, Density.cc, WellTemperedEnsemble.cc) are compiled from /root/reference, unmodified, against the HOOMD stand-in in oracle/ref_shim/
(`make -C oracle ref`, oracle/ref_capi.cc) and run on small seeded inputs.  These vectors are therefore outputs of the
reference itself (with BoxDim and kiss_fft restated by the stand-in) -- unlike golden.npz, which pins the oracle.
/root/reference only exists in the build container, so the vectors are committed.  Run from the repo root:
    python tests/golden/make_ref_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyref                     # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def postype(pos, types):
    out = np.empty((pos.shape[0], 4), dtype=np.float32)
    out[:, :3] = pos
    out[:, 3] = np.asarray(types, np.int32).view(np.float32)
    return out


def mesh_case(seed, N, dims, L, modes):
    rng = np.random.default_rng(seed)
    Lf = np.asarray(L, float)
    pos = ((rng.random((N, 3)) - 0.5) * Lf).astype(np.float32)
    pos[0] = [np.float32(Lf[0]) / 2, 0, 0]                                      # upper face: cell wraps to 0
    pos[1] = [-np.float32(Lf[0]) / 2, np.float32(Lf[1]) / 2, -np.float32(Lf[2]) / 2]
    pos[2] = np.nextafter((Lf / 2).astype(np.float32), np.float32(0))
    # particles a few ulps around cell faces: these decide whether makeFraction multiplies by Linv or divides by L
    n = np.asarray(dims)
    for i in range(3, min(N, 400)):
        d = i % 3
        k = rng.integers(0, n[d] + 1)
        x = np.float32(-Lf[d] / 2 + k * Lf[d] / n[d])
        for _ in range(int(rng.integers(0, 4))):
            x = np.nextafter(x, np.float32(1e9 if rng.random() < 0.5 else -1e9))
        pos[i, d] = x
    # HOOMD keeps particles inside the box it computes with: pull everything that float rounding pushed outside the
    # DOUBLE box [-L/2, L/2] back by whole ulps (both builds of the reference are fed the same float positions)
    for d in range(3):
        for _ in range(4):
            lo_out = pos[:, d].astype(np.float64) < -Lf[d] / 2
            hi_out = pos[:, d].astype(np.float64) > Lf[d] / 2
            pos[lo_out, d] = np.nextafter(pos[lo_out, d], np.float32(1e9))
            pos[hi_out, d] = np.nextafter(pos[hi_out, d], np.float32(-1e9))
    return postype(pos, rng.integers(0, len(modes), N))


out = {}
VIR_KMIN, VIR_KMAX, VIR_N = 0.5, 6.0, 64
out["virial_table"] = np.array([VIR_KMIN, VIR_KMAX, VIR_N])
MESH = [("m0", 11, 1500, (32, 16, 16), (10.0, 7.3, 5.1), (1.0,), 0.7),
        ("m1", 12, 3000, (32, 32, 32), (10.159366, 10.159366, 10.159366), (1.0, -1.0), -1.3),
        ("m2", 13, 2000, (64, 16, 32), (21.1, 6.0, 9.7), (1.0, -0.5, 2.0), 0.25)]
for name, seed, N, dims, L, modes, bias in MESH:
    pt = mesh_case(seed, N, dims, L, modes)
    out[name + "_postype"] = pt
    out[name + "_cfg"] = np.array(list(dims) + list(L) + [bias] + list(modes), dtype=np.float64)
    for prec in ("f64", "f32"):
        r = pyref.mesh(dims, modes, L, pt, bias, prec)
        out["%s_%s_cv" % (name, prec)] = np.array([r["cv"], r["mode_sq"]])
        out["%s_%s_force" % (name, prec)] = r["force"] if prec == "f64" else r["force"].astype(np.float32)
        out["%s_%s_rho" % (name, prec)] = r["rho"] if prec == "f64" else r["rho"].astype(np.float32)
        if prec == "f64":
            out["%s_%s_inv" % (name, prec)] = r["inv"]
            out["%s_%s_qmax" % (name, prec)] = pyref.mesh_qmax(dims, modes, L, pt, prec)
            # k-space virial with a kernel table (K is not used by the virial, its derivative dK is): dK(k) = -2 (k - 2) exp(-(k - 2)^2)
            kt = np.linspace(VIR_KMIN, VIR_KMAX, VIR_N)
            out["%s_%s_virial" % (name, prec)] = pyref.mesh_virial(dims, modes, L, pt, np.exp(-(kt - 2.0) ** 2), -2.0 * (kt - 2.0) * np.exp(-(kt - 2.0) ** 2),
                                                                  VIR_KMIN, VIR_KMAX, bias, prec)

# triclinic boxes and mesh sizes that are not powers of two: not on the device path yet, but the oracle must already be
# right about them (keys t*: cfg = dims, L, tilt, bias, modes)
TRI = [("t0", 31, 2500, (32, 16, 32), (12.0, 9.0, 10.0), (0.2, -0.1, 0.15), (1.0, -1.0), 0.8),
       ("t1", 32, 2000, (24, 20, 18), (11.0, 9.5, 8.0), (0.0, 0.0, 0.0), (1.0,), 0.8),
       ("t2", 33, 1500, (20, 12, 18), (10.0, 7.0, 9.0), (-0.15, 0.1, 0.05), (1.0, -0.5), -0.6)]
for name, seed, N, dims, L, tilt, modes, bias in TRI:
    rng = np.random.default_rng(seed)
    f = rng.random((N, 3)) * 0.998 + 0.001
    Lf = np.asarray(L)
    v = -Lf / 2 + f * Lf                                  # BoxDim::makeCoordinates: lo + f L, then the shear
    v[:, 0] += tilt[0] * v[:, 1] + tilt[1] * v[:, 2]
    v[:, 1] += tilt[2] * v[:, 2]
    pt = postype(v.astype(np.float32), rng.integers(0, len(modes), N))
    out[name + "_postype"] = pt
    out[name + "_cfg"] = np.array(list(dims) + list(L) + list(tilt) + [bias] + list(modes), dtype=np.float64)
    for prec in ("f64", "f32"):
        r = pyref.mesh(dims, modes, L, pt, bias, prec, tilt=tilt)
        out["%s_%s_cv" % (name, prec)] = np.array([r["cv"], r["mode_sq"]])
        out["%s_%s_rho" % (name, prec)] = r["rho"] if prec == "f64" else r["rho"].astype(np.float32)
        if prec == "f64":
            out["%s_%s_force" % (name, prec)] = r["force"]

LAM = [("l0", 21, 4000, (16.0, 16.0, 16.0), (0, 0, 0), [(0, 0, 3), (0, 3, 0), (3, 0, 0)], (1.0, -1.0), 0.83),
       ("l1", 22, 3000, (20.0, 24.0, 18.0), (0.1, -0.2, 0.05), [(1, 2, 0), (2, -1, 1)], (1.0, -1.0, 0.5), -0.4)]
for name, seed, N, L, tilt, lv, modes, bias in LAM:
    rng = np.random.default_rng(seed)
    pos = ((rng.random((N, 3)) - 0.5) * np.asarray(L)).astype(np.float32)
    pt = postype(pos, rng.integers(0, len(modes), N))
    out[name + "_postype"] = pt
    out[name + "_cfg"] = np.array(list(L) + list(tilt) + [bias, len(lv)] + [c for v in lv for c in v] + list(modes), dtype=np.float64)
    for prec in ("f64", "f32"):
        r = pyref.lamellar(modes, lv, L, pt, bias, prec, tilt)
        out["%s_%s_cv" % (name, prec)] = np.array([r["cv"]])
        out["%s_%s_modes" % (name, prec)] = r["modes"]
        out["%s_%s_force" % (name, prec)] = r["force"]

# umbrella: every kind, inside / outside the flat region (through CollectiveVariable::computeForces of a Lamellar CV)
pt = out["l0_postype"]
rows = []
for kind, kw in (("harmonic", dict(cv0=0.02, kappa=50.0)), ("harmonic", dict(cv0=-0.1, kappa=50.0, width_flat=0.5)),
                 ("linear", dict(cv0=0.1, scale=3.0)), ("wall", dict(cv0=0.3, kappa=1.5, scale=0.1)),
                 ("gaussian", dict(cv0=0.1, kappa=0.5, scale=2.0)), ("no_umbrella", {})):
    r = pyref.umbrella(kind, (1.0, -1.0), [(0, 0, 3), (0, 3, 0), (3, 0, 0)], (16.0, 16.0, 16.0), pt, bias_in=0.37, **kw)
    unit = pyref.lamellar((1.0, -1.0), [(0, 0, 3), (0, 3, 0), (3, 0, 0)], (16.0, 16.0, 16.0), pt, 1.0)["force"]
    i = np.argmax(np.abs(unit[:, 0]))
    rows.append([dict(no_umbrella=0, linear=1, harmonic=2, wall=3, gaussian=4)[kind], kw.get("cv0", 0.0), kw.get("kappa", 1.0),
                 kw.get("width_flat", 0.0), kw.get("scale", 1.0), r["cv"], r["energy"], r["force"][i, 0] / unit[i, 0]])
out["umbrella_rows"] = np.array(rows)       # kind, cv0, kappa, width, scale, cv, energy, bias factor that reached the force

asp = []
for d1, d2, L, tilt, bias in ((0, 1, (10.0, 12.0, 9.0), (0, 0, 0), 0.5), (2, 0, (10.0, 12.0, 9.0), (0.1, -0.2, 0.3), -1.5),
                              (1, 2, (7.0, 7.0, 21.0), (0, 0, 0), 2.0)):
    cv, vir = pyref.aspect(d1, d2, L, bias, "f64", tilt)
    asp.append([d1, d2] + list(L) + list(tilt) + [bias, cv] + list(vir))
out["aspect_rows"] = np.array(asp)

ig = []
for lengths in ((20, 30), (256, 256), (12, 9, 7), (400,)):
    n = pyref.indexgrid_num(lengths)
    rng = np.random.default_rng(len(lengths))
    for idx in rng.integers(0, n, 5):
        c = pyref.indexgrid_coords(lengths, int(idx))
        ig.append(list(lengths) + [0] * (3 - len(lengths)) + [len(lengths), n, int(idx), pyref.indexgrid_index(lengths, c)] + list(c) + [0] * (3 - len(c)))
out["indexgrid_rows"] = np.array(ig, dtype=np.int64)

# IntegratorMetaDynamics grid bias (prepRun + updateBiasPotential per step, prescribed CV values incl. values slightly
# off the grid and the forward / backward difference branches)
GRID = [("g1", dict(cv_min=[-2.0], cv_max=[2.0], num_points=[400], sigma=[0.05])),
        ("g2", dict(cv_min=[0.0, 0.0], cv_max=[1.0, 2.0], num_points=[20, 30], sigma=[0.25, 0.1])),
        ("g3", dict(cv_min=[0.0, -1.0, 2.0], cv_max=[1.0, 1.0, 3.0], num_points=[12, 9, 7], sigma=[0.2, 0.3, 0.25]))]
for name, cfg in GRID:
    d = len(cfg["num_points"])
    rng = np.random.default_rng(100 + d)
    lo, hi, npts = np.array(cfg["cv_min"]), np.array(cfg["cv_max"]), np.array(cfg["num_points"])
    s_, vals = lo + (hi - lo) * rng.random(d), []
    for t in range(16):
        s_ = np.clip(s_ + 0.05 * (hi - lo) * rng.normal(size=d), lo - 1.6 * (hi - lo) / (npts - 1), hi + 0.6 * (hi - lo) / (npts - 1))
        if t == 5:
            s_ = lo + 0.3 * (hi - lo) / (npts - 1)
        if t == 9:
            s_ = hi - 0.3 * (hi - lo) / (npts - 1)
        if t == 12:
            s_ = lo - 0.4 * (hi - lo) / (npts - 1)          # less than one spacing below the grid: first bin on x86-64
        vals.append(s_.copy())
    out[name + "_cfg"] = np.array([d] + cfg["cv_min"] + cfg["cv_max"] + cfg["num_points"] + cfg["sigma"], dtype=np.float64)
    out[name + "_vals"] = np.array(vals)
    for wt in (0, 1):
        r = pyref.grid_sequence(cfg["cv_min"], cfg["cv_max"], cfg["num_points"], cfg["sigma"], vals, list(range(16)), W=0.8, T_shift=7.0,
                                T=1.3, stride=3, well_tempered=bool(wt))
        for k in ("bias", "grid", "reweighted", "weight", "sigma_grid", "hist", "hist_gauss", "hist_delta"):
            out["%s_wt%d_%s" % (name, wt, k)] = r[k]
        out["%s_wt%d_scalars" % (name, wt)] = np.array([r["bias_potential"], r["reweight"], r["num_gaussians"]])

# adaptive Gaussians: computeSigma from prescribed per-particle gradients (all CVs differentiable; one that is not)
rng = np.random.default_rng(77)
NA = 50
cfg = dict(cv_min=[0.0, 0.0], cv_max=[1.0, 2.0], num_points=[20, 30], sigma=[0.25, 0.1])
g0 = np.abs(rng.normal(0.0, 0.05, (NA, 4))).astype(np.float32)          # positive components: the reference takes sqrt of the
g1 = np.abs(rng.normal(0.0, 0.08, (NA, 4))).astype(np.float32)          # off-diagonal sums (NaN otherwise -- restated, not tested)
out["ad_grads"] = np.stack([g0, g1])
vals = [[0.3 + 0.02 * t, 1.0 + 0.03 * t] for t in range(8)]
out["ad_vals"] = np.array(vals)
for tag, can in (("ad_all", [1, 1]), ("ad_one", [1, 0])):
    r = pyref.grid_adaptive(cfg["cv_min"], cfg["cv_max"], cfg["num_points"], cfg["sigma"], out["ad_grads"], can, 0.7, vals, list(range(8)),
                            W=0.8, T_shift=7.0, T=1.3, stride=2, well_tempered=True)
    for k in ("bias", "sigma_inv", "grid", "sigma_grid"):
        out["%s_%s" % (tag, k)] = r[k]

# WellTemperedEnsemble (CPU branch of the reference's own class): CV and the scaling of net force / torque / virial
rng = np.random.default_rng(41)
NW = 777
out["wte_force"] = rng.standard_normal((NW, 4)).astype(np.float32)
out["wte_torque"] = rng.standard_normal((NW, 4)).astype(np.float32)
out["wte_virial"] = rng.standard_normal((6, NW)).astype(np.float32)
out["wte_cfg"] = np.array([1.25, 0.5] + list(np.arange(1, 7.0)))             # external energy, bias, external virial
for prec in ("f64", "f32"):
    r = pyref.wte(out["wte_force"], out["wte_torque"], out["wte_virial"], 1.25, np.arange(1, 7.0), 0.5, prec)
    out["wte_%s_pe" % prec] = np.array([r["pe"]])
    out["wte_%s_force" % prec] = r["force"]
    out["wte_%s_torque" % prec] = r["torque"]
    out["wte_%s_virial" % prec] = r["virial"]
    out["wte_%s_external_virial" % prec] = r["external_virial"]

np.savez_compressed(os.path.join(HERE, "ref_golden.npz"), **out)
print("wrote ref_golden.npz:", {k: v.shape for k, v in out.items() if not k.endswith("postype")})
