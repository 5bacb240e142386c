"""GPU tests of the z-slab sharded mesh path: all ranks are emulated bulk-synchronously in ONE process on one GPU
(sharded.LocalComm), which exercises every slab stage and the exact buffer layouts of the collectives; the result
must match the single-plan path and the oracle.  (Real multi-process NCCL runs happen in bench.py --gpus N; the
communication pattern itself is covered on CPU with gloo in tests/test_distributed_cpu.py.)"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from metadynamics_plugin_b200 import ops, sharded
    return ops, sharded


@pytest.mark.parametrize("N,dims,L,P,modes", [
    (20000, (64, 32, 32), (20.0, 11.0, 13.0), 2, (1.0,)),
    (60000, (128, 32, 64), (40.0, 10.0, 20.0), 4, (1.0, -1.0)),
    (200000, (256, 64, 64), (64.0, 16.0, 16.0), 8, (1.0,)),
    (3000, (64, 16, 16), (9.0, 5.0, 6.0), 2, (1.0, 0.5)),          # 8^3 tiles, one tile row per slab
])
def test_mesh_slab_matches_single_plan_and_oracle(gpu, oracle, N, dims, L, P, modes):
    import torch
    ops, sharded = gpu
    nx, ny, nz = dims
    rng = np.random.default_rng(N + P)
    Lf = np.asarray(L, float)
    pos = ((rng.random((N, 3)) - 0.5) * Lf).astype(np.float32)
    pos[0] = [0.0, 0.0, np.float32(Lf[2]) / 2]                     # on the upper z face -> wraps to plane 0 (rank 0)
    pos[1, 2] = np.nextafter(np.float32(-Lf[2] / 2 + Lf[2] / P), np.float32(-1e9))     # just below a slab boundary
    types = rng.integers(0, len(modes), N).astype(np.int32)
    owner = sharded.slab_of(pos[:, 2], Lf[2], nz, P)
    box = ops.Box.make(Lf)
    ranks = [sharded.MeshSlabRank(nx, ny, nz, P, r, modes) for r in range(P)]
    for r in ranks:
        r.set(1, 1)
        r.set(3, 1)
    idx = [np.nonzero(owner == r)[0] for r in range(P)]
    pts = [ops.make_postype(pos[i], types[i]) for i in idx]
    bias = torch.tensor([0.9], dtype=torch.float64, device="cuda")
    comm = sharded.LocalComm(P)
    cvs, forces = sharded.mesh_slab_step_local(ranks, comm, pts, N, box, bias)
    cv = cvs[0].cpu().item()
    assert all(c.cpu().item() == cv for c in cvs)
    assert all(r.sums.cpu()[2].item() == 0 for r in ranks)        # no particle outside its slab
    f = np.zeros((N, 4), np.float32)
    for i, fr in zip(idx, forces):
        f[i] = fr.cpu().numpy()

    # single-plan path on the same GPU
    single = ops.Mesh(nx, ny, nz, modes)
    d_all = ops.make_postype(pos, types)
    cv1 = single.compute_cv(d_all, N, box).cpu().item()
    f1 = single.forces(d_all, N, box, bias).cpu().numpy()
    assert cv == pytest.approx(cv1, rel=2e-7)
    assert np.abs(f - f1).max() < 2e-6 * np.abs(f1).max()
    single.set(1, 1)
    single.compute_cv(d_all, N, box)
    rho = np.concatenate([r.local_mesh(1) for r in ranks], axis=0)
    fx = {r.stats()["fx_scale"] for r in ranks} | {single.stats()["fx_scale"]}
    if len(fx) == 1:                                              # equal fixed-point scales: the sharded density is bit-identical
        assert np.array_equal(rho, single.rho())

    # oracle
    h_pt = oracle.make_postype(pos, types)
    o = oracle.Mesh(nx, ny, nz, modes, Lf, N, "f64", literal_copysignf=False)
    cvo = o.current_value(h_pt)
    fo = o.forces(h_pt, 0.9)
    assert cv == pytest.approx(cvo, rel=1e-6)
    assert np.abs(f - fo).max() < 1e-5 * np.abs(fo).max()
    o32 = oracle.Mesh(nx, ny, nz, modes, Lf, N, "f32")
    o32.assign(h_pt)
    cells = np.zeros((N, 3), np.int32)
    for i, r in zip(idx, ranks):
        cells[i] = r.cells()
    assert np.array_equal(cells, o32.cells())                     # bit-exact global cell indices
    assert np.abs(rho - o.mesh).max() < 2e-6 * max(1.0, np.abs(o.mesh).max())


@pytest.mark.parametrize("N,dims,L,P,modes", [
    (20000, (64, 32, 32), (20.0, 11.0, 13.0), 2, (1.0,)),
    (60000, (128, 32, 64), (40.0, 10.0, 20.0), 4, (1.0, -1.0)),
    (200000, (256, 64, 64), (64.0, 16.0, 16.0), 8, (1.0,)),
    (1 << 21, (256, 256, 256), (128.0, 128.0, 128.0), 8, (1.0,)),       # the C4 mesh on 8 ranks: fft_y<256>, fft_z<256>, 16^3 tiles
])
def test_mesh_slab_peer_memory_path(gpu, oracle, N, dims, L, P, modes):
    """The peer-memory (P2P) slab path, all ranks emulated in one process: transposes fused into the FFT sweeps, pushed
    halos, rank-ordered reductions.  Must agree with the single plan (bitwise for the density when the scales agree) and
    with the staged path that leaves the collectives to the caller."""
    import torch
    ops, sharded = gpu
    nx, ny, nz = dims
    rng = np.random.default_rng(N + 3 * P)
    Lf = np.asarray(L, float)
    pos = ((rng.random((N, 3)) - 0.5) * Lf).astype(np.float32)
    pos[0] = [0.0, 0.0, np.float32(Lf[2]) / 2]
    types = rng.integers(0, len(modes), N).astype(np.int32)
    owner = sharded.slab_of(pos[:, 2], Lf[2], nz, P)
    box = ops.Box.make(Lf)
    idx = [np.nonzero(owner == r)[0] for r in range(P)]
    pts = [ops.make_postype(pos[i], types[i]) for i in idx]
    bias = torch.tensor([0.9], dtype=torch.float64, device="cuda")
    ranks = [sharded.MeshSlabRank(nx, ny, nz, P, r, modes) for r in range(P)]
    for r in ranks:
        r.set(1, 1)
    sharded.connect_local(ranks)
    for _ in range(2):                                            # twice: buffers and flags are reused across steps
        cvs, forces = sharded.mesh_slab_p2p_step_local(ranks, pts, N, box, bias)
    cv = cvs[0].cpu().item()
    assert all(c.cpu().item() == cv for c in cvs)
    f = np.zeros((N, 4), np.float32)
    for i, fr in zip(idx, forces):
        f[i] = fr.cpu().numpy()
    # staged path (caller-side collectives, emulated)
    ranks2 = [sharded.MeshSlabRank(nx, ny, nz, P, r, modes) for r in range(P)]
    cvs2, forces2 = sharded.mesh_slab_step_local(ranks2, sharded.LocalComm(P), pts, N, box, bias)
    assert cv == pytest.approx(cvs2[0].cpu().item(), rel=1e-12)
    f2 = np.zeros((N, 4), np.float32)
    for i, fr in zip(idx, forces2):
        f2[i] = fr.cpu().numpy()
    assert np.abs(f - f2).max() <= 1e-6 * np.abs(f2).max()
    # single plan and oracle
    single = ops.Mesh(nx, ny, nz, modes)
    single.set(1, 1)
    d_all = ops.make_postype(pos, types)
    cv1 = single.compute_cv(d_all, N, box).cpu().item()
    assert cv == pytest.approx(cv1, rel=2e-7)
    rho = np.concatenate([r.local_mesh(1) for r in ranks], axis=0)
    if len({r.stats()["fx_scale"] for r in ranks} | {single.stats()["fx_scale"]}) == 1:
        assert np.array_equal(rho, single.rho())
    h_pt = oracle.make_postype(pos, types)
    o = oracle.Mesh(nx, ny, nz, modes, Lf, N, "f64", literal_copysignf=False)
    assert cv == pytest.approx(o.current_value(h_pt), rel=1e-6)
    fo = o.forces(h_pt, 0.9)
    assert np.abs(f - fo).max() < 1e-5 * np.abs(fo).max()


@pytest.mark.parametrize("N,dims,L,tilt,P,modes,literal", [
    (20000, (64, 32, 32), (20.0, 11.0, 13.0), (0.03, -0.02, 0.04), 2, (1.0, -1.0), True),
    (60000, (128, 32, 64), (40.0, 10.0, 20.0), (0.3, 0.2, -0.25), 4, (1.0, -1.0), False),
])
def test_mesh_slab_triclinic(gpu, oracle, N, dims, L, tilt, P, modes, literal):
    """z slabs of a triclinic box (the shear leaves the z fraction alone, so the slab of a particle is still a function of
    z): staged and peer-memory paths against the single plan and the oracle."""
    import torch
    from conftest import triclinic_case
    ops, sharded = gpu
    nx, ny, nz = dims
    pos, types = triclinic_case(N, L, tilt, len(modes), N + P, faces=False)
    owner = sharded.slab_of(pos[:, 2], L[2], nz, P)
    box = ops.Box.make(L, tilt)
    idx = [np.nonzero(owner == r)[0] for r in range(P)]
    pts = [ops.make_postype(pos[i], types[i]) for i in idx]
    bias = torch.tensor([0.9], dtype=torch.float64, device="cuda")

    def make_ranks():
        rr = [sharded.MeshSlabRank(nx, ny, nz, P, r, modes) for r in range(P)]
        for r in rr:
            r.set(16, 1 if literal else 0)
        return rr

    def collect(forces):
        f = np.zeros((N, 4), np.float32)
        for i, fr in zip(idx, forces):
            f[i] = fr.cpu().numpy()
        return f

    ranks = make_ranks()
    cvs, forces = sharded.mesh_slab_step_local(ranks, sharded.LocalComm(P), pts, N, box, bias)
    cv, f = cvs[0].cpu().item(), collect(forces)
    assert all(r.sums.cpu()[2].item() == 0 for r in ranks)        # no particle outside its slab
    ranks2 = make_ranks()
    sharded.connect_local(ranks2)
    for _ in range(2):
        cvs2, forces2 = sharded.mesh_slab_p2p_step_local(ranks2, pts, N, box, bias)
    assert cvs2[0].cpu().item() == pytest.approx(cv, rel=1e-12)
    assert np.abs(collect(forces2) - f).max() <= 1e-6 * np.abs(f).max()
    single = ops.Mesh(nx, ny, nz, modes)
    single.set(16, 1 if literal else 0)
    d_all = ops.make_postype(pos, types)
    cv1 = single.compute_cv(d_all, N, box).cpu().item()
    f1 = single.forces(d_all, N, box, bias).cpu().numpy()
    assert cv == pytest.approx(cv1, rel=2e-7)
    assert np.abs(f - f1).max() < 2e-6 * np.abs(f1).max()
    h_pt = oracle.make_postype(pos, types)
    o = oracle.Mesh(nx, ny, nz, modes, L, N, "f64", tilt=tilt, literal_copysignf=False, literal_tilt_offset=literal)
    assert cv == pytest.approx(o.current_value(h_pt), rel=1e-6)
    fo = o.forces(h_pt, 0.9)
    assert np.abs(f - fo).max() < 1e-5 * np.abs(fo).max()


def test_mesh_slab_counts_misplaced_particles(gpu):
    import torch
    ops, sharded = gpu
    ranks = [sharded.MeshSlabRank(64, 16, 16, 2, r, [1.0]) for r in range(2)]
    pos = np.zeros((10, 3), np.float32)
    pos[:, 2] = -3.0                                               # all in slab 0
    ranks[1].stage_spread(ops.make_postype(pos), ops.Box.make(8.0))
    assert ranks[1].sums.cpu()[2].item() == 10                    # reported, not silently dropped


def test_slab_rejects_bad_decompositions(gpu):
    ops, sharded = gpu
    from metadynamics_plugin_b200._abi import MetadError
    with pytest.raises(MetadError):
        sharded.MeshSlabRank(64, 16, 16, 4, 0, [1.0])             # nx/2/P = 8 < 16
    with pytest.raises(MetadError):
        sharded.MeshSlabRank(128, 16, 16, 3, 0, [1.0])            # not a power of two
    with pytest.raises(MetadError, match="power of two"):
        sharded.MeshSlabRank(96, 32, 32, 2, 0, [1.0])             # like the reference under domain decomposition (OrderParameterMesh.cc:70-79)


def test_lamellar_sharded_local(gpu, oracle):
    import torch
    ops, sharded = gpu
    N, L, P = 50000, 17.0, 4
    rng = np.random.default_rng(5)
    pos = ((rng.random((N, 3)) - 0.5) * L).astype(np.float32)
    types = rng.integers(0, 2, N).astype(np.int32)
    lv, modes = [(0, 0, 3), (1, 1, 0)], [1.0, -1.0]
    box = ops.Box.make(L)
    parts = np.array_split(np.arange(N), P)
    lams = [ops.Lamellar(modes, lv) for _ in range(P)]
    pts = [ops.make_postype(pos[i], types[i]) for i in parts]
    for lam, pt in zip(lams, pts):
        lam.compute_modes(pt, N, box, finalize=False)
    sharded.LocalComm(P).all_reduce_sum([lam.modes for lam in lams])
    cvs = [lam.finalize(N).cpu().item() for lam in lams]
    cvo, _ = oracle.lamellar_cv(oracle.make_postype(pos, types), N, modes, lv, L)
    assert all(c == cvs[0] for c in cvs)
    assert abs(cvs[0] - cvo) < 2e-6 * np.sqrt(N) * 2 / N


@pytest.mark.parametrize("P,n", [(1, 6), (2, 1), (4, 6), (8, 32)])
def test_peer_allreduce_local(gpu, P, n):
    """metad_peer_allreduce_sum with all ranks in one process: rank-ordered sums, identical on every rank, table slots
    alternate with the epoch (three rounds)."""
    import torch
    ops, sharded = gpu
    grp = sharded.LocalPeerGroup(P)
    rng = np.random.default_rng(P + n)
    for rnd in range(3):
        vals = rng.normal(size=(P, n))
        ts = [torch.tensor(v, dtype=torch.float64, device="cuda") for v in vals]
        grp.all_reduce_sum(ts)
        want = np.zeros(n)
        for r in range(P):
            want = want + vals[r]                     # rank order
        for t in ts:
            np.testing.assert_array_equal(t.cpu().numpy(), want)


def test_lamellar_and_wte_sharded_over_peer_memory_local(gpu, oracle):
    """LamellarSharded / WTESharded with the peer-memory all-reduce (all ranks emulated in one process)."""
    import torch
    ops, sharded = gpu
    N, L, P = 60000, 17.0, 4
    rng = np.random.default_rng(6)
    pos = ((rng.random((N, 3)) - 0.5) * L).astype(np.float32)
    types = rng.integers(0, 2, N).astype(np.int32)
    lv, modes = [(0, 0, 3), (1, 1, 0), (2, 0, 1)], [1.0, -1.0]
    box = ops.Box.make(L)
    parts = np.array_split(np.arange(N), P)
    lams = [ops.Lamellar(modes, lv) for _ in range(P)]
    pts = [ops.make_postype(pos[i], types[i]) for i in parts]
    for lam, pt in zip(lams, pts):
        lam.compute_modes(pt, N, box, finalize=False)
    sharded.LocalPeerGroup(P).all_reduce_sum([lam.modes for lam in lams])
    cvs = [lam.finalize(N).cpu().item() for lam in lams]
    cvo, _ = oracle.lamellar_cv(oracle.make_postype(pos, types), N, modes, lv, L)
    assert all(c == cvs[0] for c in cvs)
    assert abs(cvs[0] - cvo) < 2e-6 * np.sqrt(N) * 2 / N
    nf = rng.normal(size=(N, 4)).astype(np.float32)
    pes = [ops.wte_reduce(torch.from_numpy(nf[i]).cuda(), 0.0) for i in parts]
    sharded.LocalPeerGroup(P).all_reduce_sum(pes)
    assert all(p.cpu().item() == pes[0].cpu().item() for p in pes)
    assert pes[0].cpu().item() + 1.25 == pytest.approx(oracle.wte_pe(nf, 1.25), rel=1e-12)
