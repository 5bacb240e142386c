"""World-size-2 (and 4) gloo tests of the multi-GPU drivers in metadynamics_plugin_b200/sharded.py, on CPU.

The drivers (MeshSlab, LamellarSharded) are written against a small communicator interface; here they run over real
torch.distributed process groups (gloo) with a numpy stand-in for the per-rank compute stages that honours the exact
buffer layouts of the C ABI (include/metad_b200.h: halo messages, the [dest][plane][y][kx] all-to-all packing, the
packed kx = 0 column, the kx pencil).  What is tested is the communication pattern: which buffer goes to which rank,
in which order, and that the assembled result equals the oracle's single-domain answer.
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


# ---------------------------------------------------------------------------------------------- numpy stand-in
def tsc_w(s):
    return np.stack([0.5 * (0.5 - s) ** 2, 0.75 - s * s, 0.5 * (0.5 + s) ** 2], -1)


def tsc_d(s):
    return np.stack([s - 0.5, -2.0 * s, s + 0.5], -1)


def cells_and_shifts(pos, L, dims):
    n = np.asarray(dims)
    Lf = np.asarray(L, np.float32)
    f = (pos.astype(np.float32) + Lf / np.float32(2)) / Lf
    c = (f * n.astype(np.float32)).astype(np.int64)
    c[c == n] = 0
    s = (pos.astype(np.float64) + np.asarray(L) / 2) * (n / np.asarray(L)) - (c + 0.5)
    s = np.where(s > n / 2, s - n, s)
    return c, s


class NumpySlabRank:
    """Same attributes and stage methods as sharded.MeshSlabRank, computed with numpy on CPU tensors."""
    FX = 2.0 ** 20

    def __init__(self, nx, ny, nz, n_ranks, rank, mode):
        import torch
        self.torch = torch
        self.dims, self.P, self.rank = (nx, ny, nz), n_ranks, rank
        self.nzl, self.kxl = nz // n_ranks, nx // 2 // n_ranks
        self.mode = np.asarray(mode, float)
        m_local = nx * ny * self.nzl
        self.sums = torch.zeros(3, dtype=torch.float64)
        self.ghost_send = torch.zeros(2, ny * nx + 4, dtype=torch.int32)
        self.ghost_recv = torch.zeros(2, ny * nx + 4, dtype=torch.int32)
        self.send = torch.zeros(m_local, dtype=torch.float32)
        self.pencil = torch.zeros(m_local, dtype=torch.float32)
        self.cv = torch.zeros(1, dtype=torch.float64)
        self.inv_send = torch.zeros(2, ny, nx, dtype=torch.float32)
        self.inv_ghost = torch.zeros(2, ny, nx, dtype=torch.float32)

    def stage_spread(self, postype, box):
        nx, ny, nz = self.dims
        pt = postype.numpy()
        self.L = np.array([box.L[0], box.L[1], box.L[2]])
        a = self.mode[pt[:, 3].view(np.int32)]
        c, s = cells_and_shifts(pt[:, :3], self.L, self.dims)
        zl = c[:, 2] - self.rank * self.nzl
        assert ((zl >= 0) & (zl < self.nzl)).all()
        rho = np.zeros((self.nzl + 2, ny, nx), np.int64)           # plane 0 / nzl+1 = ghost planes
        wx, wy, wz = tsc_w(s[:, 0]), tsc_w(s[:, 1]), tsc_w(s[:, 2])
        for k in range(3):
            for j in range(3):
                for i in range(3):
                    v = np.rint(a * wx[:, i] * wy[:, j] * wz[:, k] * self.FX).astype(np.int64)
                    np.add.at(rho, (zl + k, (c[:, 1] + j - 1) % ny, (c[:, 0] + i - 1) % nx), v)
        self.rho_i = rho[1:-1].copy()
        gs = self.ghost_send.numpy()
        gs[0, :ny * nx] = rho[0].reshape(-1)
        gs[1, :ny * nx] = rho[-1].reshape(-1)
        gs[:, ny * nx] = np.float32(1.0 / self.FX).view(np.int32)
        self.sums[0], self.sums[1], self.sums[2] = float((a * a).sum()), float(a.sum()), 0.0

    def stage_fft_x(self):
        nx, ny, nz = self.dims
        gr = self.ghost_recv.numpy()
        assert gr[0, ny * nx] == np.float32(1.0 / self.FX).view(np.int32)     # the neighbour's scale travelled with the plane
        rho = self.rho_i.copy()
        rho[0] += gr[0, :ny * nx].reshape(ny, nx)
        rho[-1] += gr[1, :ny * nx].reshape(ny, nx)
        self.rho = rho / self.FX
        r = self.rho - self.sums[1].item() / (nx * ny * nz)
        F = np.fft.rfft(r, axis=2)                                  # [zl][y][nx/2+1]
        Pk = F[:, :, :nx // 2].copy()
        Pk[:, :, 0] = F[:, :, 0].real + 1j * F[:, :, nx // 2].real  # packed slot: X0 + i X_{nx/2}
        # [dest][plane][y][kx in pencil]
        out = Pk.reshape(self.nzl, ny, self.P, self.kxl).transpose(2, 0, 1, 3)
        self.send.numpy().view(np.complex64)[:] = out.reshape(-1).astype(np.complex64)

    def stage_fft_yz(self, n_global):
        nx, ny, nz = self.dims
        Z = self.pencil.numpy().view(np.complex64).reshape(nz, ny, self.kxl).astype(np.complex128)
        Z = np.fft.fft(np.fft.fft(Z, axis=1), axis=0)
        d = 0.5 * self.sums[0].item() / n_global / n_global
        nn = lambda n: np.arange(n) < (n // 2 + n % 2)
        chi_yz = nn(nz)[:, None] & nn(ny)[None, :]
        e = 0.0
        G = np.empty_like(Z)
        for lk in range(self.kxl):
            kx = self.rank * self.kxl + lk
            if kx == 0:
                Zm = np.conj(np.roll(np.roll(Z[::-1, ::-1, 0], 1, 0), 1, 1))     # conj Z(-ky,-kz)
                A, B = (Z[:, :, 0] + Zm) / 2, (Z[:, :, 0] - Zm) / 2j
                chim = np.roll(np.roll(chi_yz[::-1, ::-1], 1, 0), 1, 1)
                chis = 0.5 * (chi_yz.astype(float) + chim)
                fa, fb = A / n_global, B / n_global
                va, vb = np.abs(fa) ** 2, np.abs(fb) ** 2
                ea = va * (va - 2 * d * chis)
                ea[0, 0] = 0.0
                e += ea.sum() + (vb * vb).sum()
                G[:, :, 0] = fa * (va - d * chis) + 1j * (fb * vb)
            else:
                f = Z[:, :, lk] / n_global
                v = np.abs(f) ** 2
                chis = 0.5 * chi_yz
                e += 2.0 * (v * (v - 2 * d * chis)).sum()
                G[:, :, lk] = f * (v - d * chis)
        self.cv[0] = 0.5 * e
        out = np.fft.ifft(np.fft.ifft(G, axis=0), axis=1) * (nz * ny)
        self.pencil.numpy().view(np.complex64)[:] = out.reshape(-1).astype(np.complex64)

    def stage_fft_x_inv(self):
        nx, ny, nz = self.dims
        R = self.send.numpy().view(np.complex64).reshape(self.P, self.nzl, ny, self.kxl).astype(np.complex128)
        Pk = R.transpose(1, 2, 0, 3).reshape(self.nzl, ny, nx // 2)
        F = np.zeros((self.nzl, ny, nx // 2 + 1), complex)
        F[:, :, 1:nx // 2] = Pk[:, :, 1:]
        F[:, :, 0] = Pk[:, :, 0].real
        F[:, :, nx // 2] = Pk[:, :, 0].imag
        self.inv = np.fft.irfft(F, n=nx, axis=2) * nx
        self.inv_send[0] = self.torch.from_numpy(self.inv[0].astype(np.float32))
        self.inv_send[1] = self.torch.from_numpy(self.inv[-1].astype(np.float32))

    def stage_forces(self, postype, n_global, box, bias, out=None):
        nx, ny, nz = self.dims
        pt = postype.numpy()
        a = self.mode[pt[:, 3].view(np.int32)]
        c, s = cells_and_shifts(pt[:, :3], self.L, self.dims)
        zl = c[:, 2] - self.rank * self.nzl
        inv = np.concatenate([self.inv_ghost[0].numpy()[None].astype(float), self.inv, self.inv_ghost[1].numpy()[None].astype(float)], 0)
        w = [tsc_w(s[:, d]) for d in range(3)]
        dw = [tsc_d(s[:, d]) for d in range(3)]
        S = np.zeros((pt.shape[0], 3))
        for k in range(3):
            for j in range(3):
                for i in range(3):
                    v = inv[zl + k, (c[:, 1] + j - 1) % ny, (c[:, 0] + i - 1) % nx]
                    S[:, 0] += dw[0][:, i] * w[1][:, j] * w[2][:, k] * v
                    S[:, 1] += w[0][:, i] * dw[1][:, j] * w[2][:, k] * v
                    S[:, 2] += w[0][:, i] * w[1][:, j] * dw[2][:, k] * v
        f = np.zeros((pt.shape[0], 4), np.float32)
        f[:, :3] = (-a * 2.0 / n_global * float(bias[0]))[:, None] * S * (np.array(self.dims) / self.L)
        return self.torch.from_numpy(f)


# ---------------------------------------------------------------------------------------------- workers
def _mesh_worker(rank, world, port, dims, L, modes, N, out_q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from metadynamics_plugin_b200 import sharded
        from metadynamics_plugin_b200._abi import Box
        rng = np.random.default_rng(123)
        pos = ((rng.random((N, 3)) - 0.5) * np.asarray(L)).astype(np.float32)
        types = rng.integers(0, len(modes), N).astype(np.int32)
        pt = np.empty((N, 4), np.float32)
        pt[:, :3] = pos
        pt[:, 3] = types.view(np.float32)
        owner = sharded.slab_of(pos[:, 2], L[2], dims[2], world)
        mine = np.nonzero(owner == rank)[0]
        comm = sharded.TorchComm()
        slab = sharded.MeshSlab(comm, *dims, modes, rank_factory=NumpySlabRank)
        box = Box.make(L)
        local = torch.from_numpy(pt[mine].copy())
        cv = slab.compute_cv(local, N, box)
        bias = torch.tensor([0.7], dtype=torch.float64)
        f = slab.forces(local, N, box, bias)
        out_q.put((rank, float(cv[0]), mine, f.numpy()))
    finally:
        dist.destroy_process_group()


def _lamellar_worker(rank, world, port, N, out_q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from metadynamics_plugin_b200 import sharded

        class NumpyLamellar:                       # stand-in for ops.Lamellar with the same attributes
            def __init__(self, mode, lv):
                self.mode, self.lv = np.asarray(mode, float), np.asarray(lv, float)
                self.modes = torch.zeros(2 * len(lv), dtype=torch.float64)
                self.cv = torch.zeros(1, dtype=torch.float64)

            def compute_modes(self, postype, n_global, box, finalize=True):
                pt = postype.numpy()
                a = self.mode[pt[:, 3].view(np.int32)]
                q = 2 * np.pi * self.lv / np.array([box.L[0], box.L[1], box.L[2]])
                ph = pt[:, :3].astype(float) @ q.T
                m = np.stack([(a[:, None] * np.cos(ph)).sum(0), (a[:, None] * np.sin(ph)).sum(0)], -1)
                self.modes[:] = torch.from_numpy(m.reshape(-1))

            def finalize(self, n_global):
                self.cv[0] = self.modes[0::2].sum() / n_global
                return self.cv

        from metadynamics_plugin_b200._abi import Box
        rng = np.random.default_rng(7)
        L = 9.0
        pos = ((rng.random((N, 3)) - 0.5) * L).astype(np.float32)
        types = rng.integers(0, 2, N).astype(np.int32)
        pt = np.empty((N, 4), np.float32)
        pt[:, :3] = pos
        pt[:, 3] = types.view(np.float32)
        part = np.array_split(np.arange(N), world)[rank]
        lv, modes = [(0, 0, 2), (1, 1, 0)], [1.0, -1.0]
        ls = sharded.LamellarSharded(sharded.TorchComm(), modes, lv, lamellar_factory=NumpyLamellar)
        cv = ls.compute_cv(torch.from_numpy(pt[part].copy()), N, Box.make(L))
        out_q.put((rank, float(cv[0])))
    finally:
        dist.destroy_process_group()


def _wte_worker(rank, world, port, N, out_q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from metadynamics_plugin_b200 import sharded

        def np_reduce(net_force, external_energy=0.0, out=None):      # stand-in for ops.wte_reduce
            if out is None:
                out = torch.zeros(1, dtype=torch.float64)
            out[0] = float(net_force.numpy()[:, 3].astype(np.float64).sum()) + external_energy
            return out

        def np_scale(nf, tq, vir, pitch, bias):                        # stand-in for ops.wte_scale
            fac = 1.0 + float(bias[0])
            nf[:, :3] *= fac
            tq[:, :3] *= fac
            vir *= fac

        rng = np.random.default_rng(99)
        nf = rng.standard_normal((N, 4)).astype(np.float32)
        part = np.array_split(np.arange(N), world)[rank]
        w = sharded.WTESharded(sharded.TorchComm(), np_reduce, np_scale)
        local = torch.from_numpy(nf[part].copy())
        cv = w.compute_cv(local, external_energy=1.25)
        cv2 = float(w.compute_cv(local, external_energy=1.25)[0])      # the cached output tensor must not accumulate
        tq = torch.zeros(len(part), 4)
        vir = torch.zeros(6, len(part))
        w.scale(local, tq, vir, len(part), torch.tensor([0.5], dtype=torch.float64))
        out_q.put((rank, float(cv[0]), cv2, part, local.numpy()))
    finally:
        dist.destroy_process_group()


class OracleWalkerGrid:
    """CPU stand-in with the interface of ops.BiasGrid's walker methods, on the oracle's restated updateBiasPotential."""

    def __init__(self, po, torch, cfg, kw):
        self.o, self.torch, self.stride = po.Grid(**cfg, **kw), torch, kw["stride"]

    def step_deposit(self, t, cv):
        self.o.update_deposit(t, cv)

    def is_deposit_step(self, t):
        return t % self.stride == 0

    def deltas_export(self):
        d = self.o.get_deltas()
        return self.torch.from_numpy(np.ascontiguousarray(d[:2]).reshape(-1)), self.torch.from_numpy(np.ascontiguousarray(d[2:]).astype(np.int32).reshape(-1))

    def deltas_import(self, dd, du):
        G = self.o.G
        self.o.set_deltas(np.concatenate([dd.numpy().reshape(2, G), du.numpy().astype(np.float64).reshape(2, G)]))

    def step_merge(self, t, cv):
        return self.o.update_merge(t, cv)


WALKER_CFG = dict(cv_min=[0.0, 0.0], cv_max=[1.0, 2.0], num_points=[20, 30], sigma=[0.25, 0.1])
WALKER_KW = dict(W=0.8, T_shift=7.0, T=1.3, stride=3, well_tempered=True)


def _walker_trajectory(world, steps=10):
    rng = np.random.default_rng(5)
    lo, hi = np.array(WALKER_CFG["cv_min"]), np.array(WALKER_CFG["cv_max"])
    s = lo + (hi - lo) * rng.random((world, 2))
    out = []
    for _ in range(steps):
        s = np.clip(s + 0.05 * (hi - lo) * rng.normal(size=(world, 2)), lo, hi - 1e-9)
        out.append(s.copy())
    return out


def _walker_worker(rank, world, port, out_q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from metadynamics_plugin_b200 import sharded
        from oracle import pyoracle as po
        wb = sharded.WalkerBias(sharded.TorchComm(), OracleWalkerGrid(po, torch, WALKER_CFG, WALKER_KW))
        bias = [np.array(wb.step(t, s[rank])) for t, s in enumerate(_walker_trajectory(world))]
        out_q.put((rank, np.array(bias), wb.grid.o.get("grid"), wb.grid.o.get("hist"), wb.grid.o.get("reweighted")))
    finally:
        dist.destroy_process_group()


def _run(worker, world, *args):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=worker, args=(r, world, port) + args + (q,)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return sorted(res, key=lambda t: t[0])


# ---------------------------------------------------------------------------------------------- tests
@pytest.mark.parametrize("world,dims,L", [(2, (64, 16, 16), (9.0, 5.0, 6.0)), (4, (128, 16, 32), (12.0, 5.0, 9.0))])
def test_mesh_slab_driver_over_gloo(oracle, world, dims, L):
    N, modes = 3000, (1.0, -0.5)
    res = _run(_mesh_worker, world, dims, L, modes, N)
    rng = np.random.default_rng(123)
    pos = ((rng.random((N, 3)) - 0.5) * np.asarray(L)).astype(np.float32)
    types = rng.integers(0, len(modes), N).astype(np.int32)
    h_pt = oracle.make_postype(pos, types)
    o = oracle.Mesh(*dims, modes, L, N, "f64", literal_copysignf=False)
    cvo = o.current_value(h_pt)
    fo = o.forces(h_pt, 0.7)
    f = np.zeros((N, 4))
    for rank, cv, mine, fr in res:
        assert cv == pytest.approx(cvo, rel=2e-5)           # complex64 transport buffers
        f[mine] = fr
    assert all(r[1] == res[0][1] for r in res)              # every rank holds the same CV after the all-reduce
    assert np.abs(f - fo).max() < 2e-4 * np.abs(fo).max()


def test_lamellar_sharded_driver_over_gloo(oracle):
    N = 4000
    res = _run(_lamellar_worker, 2, N)
    rng = np.random.default_rng(7)
    pos = ((rng.random((N, 3)) - 0.5) * 9.0).astype(np.float32)
    types = rng.integers(0, 2, N).astype(np.int32)
    cvo, _ = oracle.lamellar_cv(oracle.make_postype(pos, types), N, [1.0, -1.0], [(0, 0, 2), (1, 1, 0)], 9.0)
    assert res[0][1] == res[1][1]
    assert res[0][1] == pytest.approx(cvo, abs=1e-9)


@pytest.mark.parametrize("world", [2, 4])
def test_wte_sharded_driver_over_gloo(oracle, world):
    """WellTemperedEnsemble over particle shards: one all-reduce of a double, external energy added once."""
    N = 5000
    res = _run(_wte_worker, world, N)
    nf = np.random.default_rng(99).standard_normal((N, 4)).astype(np.float32)
    pe = oracle.wte_pe(nf, 1.25)
    for rank, cv, cv2, part, scaled in res:
        assert cv == pytest.approx(pe, rel=1e-12) and cv2 == cv
        np.testing.assert_allclose(scaled[:, :3], nf[part][:, :3] * np.float32(1.5), rtol=1e-6)
        np.testing.assert_array_equal(scaled[:, 3], nf[part][:, 3])


@pytest.mark.parametrize("world", [2, 3])
def test_multiple_walkers_over_gloo(oracle, world):
    """Multiple walkers, one process each (the reference's partitions): sharded.WalkerBias over a gloo group against the
    serial statement of IntegratorMetaDynamics.cc:392-410 (all walkers in one process, deltas summed by hand)."""
    res = _run(_walker_worker, world)
    serial = [oracle.Grid(**WALKER_CFG, **WALKER_KW) for _ in range(world)]
    bias = [[] for _ in range(world)]
    for t, s in enumerate(_walker_trajectory(world)):
        for k in range(world):
            serial[k].update_deposit(t, s[k])
        if t % WALKER_KW["stride"] == 0:
            tot = sum(o.get_deltas() for o in serial)
            for o in serial:
                o.set_deltas(tot)
        for k in range(world):
            bias[k].append(serial[k].update_merge(t, s[k]))
    for rank, b, grid, hist, rew in res:
        np.testing.assert_allclose(b, np.array(bias[rank]), rtol=1e-12, atol=1e-300)
        np.testing.assert_allclose(grid, serial[rank].get("grid"), rtol=1e-12, atol=1e-300)
        np.testing.assert_array_equal(hist, serial[rank].get("hist"))
        np.testing.assert_allclose(rew, serial[rank].get("reweighted"), rtol=1e-12)
    np.testing.assert_array_equal(res[0][2], res[1][2])        # every walker holds the same bias potential
