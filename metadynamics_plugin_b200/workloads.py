"""Synthetic workloads of BASELINE.json / SURVEY.md 8(d), as seeded numpy generators (no torch, no CUDA).

Every generator returns a dict with the HOOMD-layout particle array `postype` (float32 (N,4), type id as raw
bits in column 3), the cubic box length `L`, the per-type mode coefficients and the CV parameters.
"""
import numpy as np


def _postype(pos, types):
    out = np.empty((pos.shape[0], 4), dtype=np.float32)
    out[:, :3] = pos
    out[:, 3] = np.asarray(types, dtype=np.int32).view(np.float32)
    return out


def _wrap(pos, L):
    return ((pos + L / 2.0) % L) - L / 2.0


def _cell_order(pos, L, n):
    """Row-major cell order (z slowest): the steady-state spatially sorted order an MD engine keeps."""
    f = np.floor((pos.astype(np.float64) + L / 2.0) / L * n).astype(np.int64) % n
    key = f[:, 0] + n * (f[:, 1] + n * f[:, 2])
    return np.argsort(key, kind="stable")


def _morton_order(pos, L, n, coarse=4, seed=1):
    """Space-filling-curve order on COARSE cells (coarse^3 mesh cells each), random inside a coarse cell: what an MD engine's
    periodic SFC sort looks like some steps after the sort -- spatially local, but not sorted by mesh cell."""
    rng = np.random.default_rng(seed)
    shuffle = rng.permutation(pos.shape[0])
    c = (np.floor((pos[shuffle].astype(np.float64) + L / 2.0) / L * n).astype(np.int64) % n) // coarse
    key = np.zeros(pos.shape[0], dtype=np.int64)
    for b in range(10):
        for d in range(3):
            key |= ((c[:, d] >> b) & 1) << (3 * b + d)
    return shuffle[np.argsort(key, kind="stable")]


def c1(seed=20260101):
    """test/test_mesh.py geometry: N=1000 jittered simple-cubic lattice, L=10, mesh 32^3, harmonic umbrella."""
    rng = np.random.default_rng(seed)
    g = np.arange(10) + 0.5 - 5.0
    pos = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3) + rng.normal(0.0, 0.1, (1000, 3))
    pos = _wrap(pos, 10.0)
    cv0 = 0.025
    return dict(name="C1", kind="mesh", postype=_postype(pos, np.zeros(1000, np.int32)), L=10.0, mode=[1.0],
                mesh=(32, 32, 32), umbrella=dict(kind="harmonic", cv0=cv0, kappa=10000.0 / cv0 ** 2))


def diblock(N, L, n_period, seed):
    """Synthetic lamellar A/B melt: uniform positions, type A with probability (1 + 0.8 cos(2 pi n z/L))/2."""
    rng = np.random.default_rng(seed)
    pos = (rng.random((N, 3)) - 0.5) * L
    p_a = 0.5 * (1.0 + 0.8 * np.cos(2.0 * np.pi * n_period * pos[:, 2] / L))
    types = (rng.random(N) >= p_a).astype(np.int32)      # 0 = A, 1 = B
    return pos, types


def c2(seed=20260102, N=262144):
    L = 64.0 * (N / 262144.0) ** (1.0 / 3.0)
    pos, types = diblock(N, L, 3, seed)
    return dict(name="C2", kind="lamellar", postype=_postype(pos, types), L=L, mode=[1.0, -1.0],
                lattice_vectors=[(0, 0, 3), (0, 3, 0), (3, 0, 0)],
                grid=dict(cv_min=[-2.0], cv_max=[2.0], num_points=[400], sigma=[0.05]),
                W=1.0, deltaT=7.0, T=1.0, stride=100)


def random_mesh(N, nmesh, seed, sort=True, name="C3"):
    """sort: True / "cell" = row-major mesh-cell order (headline), "sfc" = Morton order on 4^3-cell blocks, False = random."""
    rng = np.random.default_rng(seed)
    L = float(N) ** (1.0 / 3.0)
    pos = ((rng.random((N, 3)) - 0.5) * L).astype(np.float32)
    if sort == "sfc":
        pos = pos[_morton_order(pos, L, nmesh)]
    elif sort:
        pos = pos[_cell_order(pos, L, nmesh)]
    return dict(name=name, kind="mesh", postype=_postype(pos, np.zeros(N, np.int32)), L=L, mode=[1.0],
                mesh=(nmesh, nmesh, nmesh), stride=100)


def c3(seed=20260103, N=1 << 20, sort=True):
    return random_mesh(N, 128, seed, sort, "C3")


def c4(seed=20260104, N=1 << 24, sort=True):
    return random_mesh(N, 256, seed, sort, "C4")


def c5(seed=20260105, N=1 << 23):
    L = float(N) ** (1.0 / 3.0)
    pos, types = diblock(N, L, 10, seed)
    return dict(name="C5", kind="lamellar", postype=_postype(pos, types), L=L, mode=[1.0, -1.0],
                lattice_vectors=[(0, 0, 10), (0, 10, 0), (10, 0, 0)],
                grid=dict(cv_min=[-2.0, 0.0], cv_max=[2.0, 2.0], num_points=[256, 256], sigma=[0.05, 0.1]),
                aspect=(0, 1), W=1.0, deltaT=7.0, T=1.0, stride=100)


def wte(seed=20260106, N=1 << 23):
    """WellTemperedEnsemble at the size of C5: a synthetic net-force array (force xyz, per-particle potential energy in w)."""
    rng = np.random.default_rng(seed)
    nf = rng.standard_normal((N, 4), dtype=np.float32)
    nf[:, 3] = nf[:, 3] * 0.5 - 3.0
    return dict(name="WTE", kind="wte", net_force=nf, postype=nf, stride=100)
