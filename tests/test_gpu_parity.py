"""GPU parity tests: the sm_100a path (through the C ABI, include/metad_b200.h) against the CPU oracle.

Tolerances are the north star's (BASELINE.json): CV values 1e-6 relative, forces 1e-5 (relative to max |F|),
cell / bin / grid indices bit-exact against the single-precision CPU build.  Oracle truth = the double
instance (for the mesh forces with |x| evaluated exactly, see oracle/metad_oracle.hpp on copysignf).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from metadynamics_plugin_b200 import ops, _abi       # raises if libmetad_b200.so is missing: no fallback
    return ops


def rand_pt(N, L, ntypes, seed):
    rng = np.random.default_rng(seed)
    L = np.broadcast_to(np.asarray(L, float), (3,))
    pos = ((rng.random((N, 3)) - 0.5) * L).astype(np.float32)
    return pos, rng.integers(0, ntypes, N).astype(np.int32)


def to_dev(gpu, pos, types):
    return gpu.make_postype(pos, types)


def host_pt(oracle, pos, types):
    return oracle.make_postype(pos, types)


# ------------------------------------------------------------------------------------------------ lamellar
@pytest.mark.parametrize("N,L,tilt,lv,modes", [
    (4096, 16.0, (0, 0, 0), [(0, 0, 3), (0, 3, 0), (3, 0, 0)], [1.0, -1.0]),
    (100003, (20.0, 24.0, 18.0), (0, 0, 0), [(0, 0, 5)], [1.0, -1.0, 0.5]),
    (50000, (20.0, 24.0, 18.0), (0.1, -0.2, 0.05), [(1, 2, 0), (2, -1, 1)], [1.0, -1.0]),
    (20000, 12.0, (0, 0, 0), [(i % 3, (i + 1) % 4, i % 5 + 1) for i in range(11)], [1.0, -1.0]),      # > 8 modes: two passes
    (1, 5.0, (0, 0, 0), [(0, 0, 1)], [2.0]),
])
def test_lamellar_cv_and_forces(gpu, oracle, N, L, tilt, lv, modes):
    import torch
    pos, types = rand_pt(N, L, len(modes), 7)
    d_pt = to_dev(gpu, pos, types)
    box = gpu.Box.make(np.broadcast_to(np.asarray(L, float), (3,)), tilt)
    lam = gpu.Lamellar(modes, lv)
    cv = lam.compute_modes(d_pt, N, box).cpu().item()
    fm = lam.modes.cpu().numpy().reshape(-1, 2)
    h_pt = host_pt(oracle, pos, types)
    cvo, fmo = oracle.lamellar_cv(h_pt, N, modes, lv, L, "f64", tilt)
    scale = np.sqrt(N) * max(abs(m) for m in modes)
    assert np.abs(fm - fmo).max() < 2e-6 * scale
    assert abs(cv - cvo) < max(1e-6 * abs(cvo), 2e-6 * scale * len(lv) / N)
    bias = torch.tensor([0.83], dtype=torch.float64, device="cuda")
    f = lam.forces(d_pt, N, box, bias).cpu().numpy()
    fo = oracle.lamellar_forces(h_pt, N, modes, lv, L, 0.83, "f64", tilt)
    assert np.abs(f - fo).max() < 1e-5 * np.abs(fo).max()
    assert np.all(f[:, 3] == 0)


def test_lamellar_ordered_cv_relative(gpu, oracle):
    """Lamellar-ordered melt (config C2 at reduced N): CV is O(0.1), 1e-6 relative."""
    from metadynamics_plugin_b200 import workloads
    w = workloads.c2(N=65536)
    N = w["postype"].shape[0]
    import torch
    d_pt = torch.from_numpy(w["postype"]).cuda()
    lam = gpu.Lamellar(w["mode"], w["lattice_vectors"])
    cv = lam.compute_modes(d_pt, N, gpu.Box.make(w["L"])).cpu().item()
    cvo, _ = oracle.lamellar_cv(w["postype"], N, w["mode"], w["lattice_vectors"], w["L"])
    assert abs(cvo) > 0.05
    assert cv == pytest.approx(cvo, rel=1e-6)


def test_lamellar_sharded_modes_allreduce_equivalence(gpu, oracle):
    """Particle-sharded evaluation (the multi-GPU decomposition): partial modes of the shards add up to the
    single-shard modes; finalize after the (emulated) all-reduce gives the same CV."""
    pos, types = rand_pt(30000, 14.0, 2, 3)
    lv, modes = [(0, 0, 3), (2, 1, 0)], [1.0, -1.0]
    box = gpu.Box.make(14.0)
    full = gpu.Lamellar(modes, lv)
    cv_full = full.compute_modes(to_dev(gpu, pos, types), 30000, box).cpu().item()
    acc = None
    for lo, hi in ((0, 7000), (7000, 7001), (7001, 30000)):
        part = gpu.Lamellar(modes, lv)
        part.compute_modes(to_dev(gpu, pos[lo:hi], types[lo:hi]), 30000, box, finalize=False)
        acc = part.modes.clone() if acc is None else acc + part.modes
    full.modes.copy_(acc)
    # terms are summed in fp32 over runs of 8 particles, the runs in fp64: regrouping the particles moves the sum by
    # ~1e-7/sqrt(N) of the mode amplitude (this CV is ~0 for a disordered system, so the bound is absolute)
    assert abs(full.finalize(30000).cpu().item() - cv_full) < 1e-8 * np.sqrt(30000) / 30000


# ------------------------------------------------------------------------------------------------ mesh
MESH_CASES = [
    (1000, (32, 32, 32), 10.0, (1.0,), True),
    (5000, (32, 16, 64), (10.0, 7.3, 21.1), (1.0, -1.0), True),
    (5000, (64, 32, 16), (10.0, 7.3, 21.1), (1.0, -0.5, 2.0), True),
    (131072, (64, 64, 64), 50.8, (1.0,), False),
    (300000, (128, 64, 128), (70.0, 35.0, 70.0), (1.0,), False),        # 16^3 tiles
    (3, (32, 32, 32), 6.0, (1.0,), False),
]


@pytest.mark.parametrize("N,dims,L,modes,edge", MESH_CASES)
def test_mesh_cv_forces_cells(gpu, oracle, N, dims, L, modes, edge):
    import torch
    Lf = np.broadcast_to(np.asarray(L, float), (3,))
    pos, types = rand_pt(N, Lf, len(modes), N % 1000 + 1)
    if edge:
        pos[0] = [np.float32(Lf[0]) / 2, 0, 0]
        pos[1] = [-np.float32(Lf[0]) / 2, np.float32(Lf[1]) / 2, -np.float32(Lf[2]) / 2]
        pos[2] = np.nextafter((Lf / 2).astype(np.float32), np.float32(0))
    d_pt = to_dev(gpu, pos, types)
    h_pt = host_pt(oracle, pos, types)
    box = gpu.Box.make(Lf)
    mesh = gpu.Mesh(*dims, modes)
    mesh.set(1, 1)                                               # keep rho for inspection
    mesh.set(3, 1)                                               # keep the cell indices computed by the spread
    cv = mesh.compute_cv(d_pt, N, box).cpu().item()
    m = oracle.Mesh(*dims, modes, Lf, N, "f64", literal_copysignf=False)
    cvo = m.current_value(h_pt)
    m32 = oracle.Mesh(*dims, modes, Lf, N, "f32")
    m32.assign(h_pt)
    assert np.array_equal(mesh.cells(), m32.cells())             # bit-exact cell indices
    assert mesh.mode_sq() == m.mode_sq()
    assert np.abs(mesh.rho() - m.mesh).max() < 2e-6 * max(1.0, np.abs(m.mesh).max())
    assert cv == pytest.approx(cvo, rel=1e-6)
    bias = torch.tensor([0.61], dtype=torch.float64, device="cuda")
    f = mesh.forces(d_pt, N, box, bias).cpu().numpy()
    fo = m.forces(h_pt, 0.61)
    assert np.abs(f - fo).max() < 1e-5 * np.abs(fo).max()
    assert np.all(f[:, 3] == 0)
    # determinism: a second evaluation is bitwise identical (integer density accumulation, fixed summation orders)
    cv2 = mesh.compute_cv(d_pt, N, box).cpu().item()
    f2 = mesh.forces(d_pt, N, box, bias).cpu().numpy()
    assert cv2 == cv and np.array_equal(f, f2)
    st = mesh.stats()
    assert st["rebuilds"] == 1 and st["drifted"] == 0 and st["outside_slab"] == 0 and st["range_warnings"] == 0


def _ref_gold():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_golden.npz"))


@pytest.mark.parametrize("name", ["m0", "m1", "m2"])
def test_mesh_against_reference_vectors(gpu, oracle, name):
    """The device path against outputs of the REFERENCE's own OrderParameterMesh.cc (tests/golden/ref_golden.npz, generated
    by compiling the reference's sources against a HOOMD stand-in; inputs include particles a few ulps around cell faces)."""
    import torch
    G = _ref_gold()
    c = G[name + "_cfg"]
    dims, L, bias, modes = tuple(int(v) for v in c[:3]), tuple(c[3:6]), float(c[6]), tuple(c[7:])
    pt = G[name + "_postype"]
    N = pt.shape[0]
    d_pt = torch.from_numpy(pt).cuda()
    box = gpu.Box.make(L)
    mesh = gpu.Mesh(*dims, modes)
    mesh.set(1, 1)
    cv = mesh.compute_cv(d_pt, N, box).cpu().item()
    ref_cv, ref_msq = G[name + "_f64_cv"]
    assert mesh.mode_sq() == ref_msq
    rho = G[name + "_f64_rho"]
    assert np.abs(mesh.rho() - rho).max() < 2e-6 * max(1.0, np.abs(rho).max())      # every particle in the reference's cell
    assert cv == pytest.approx(ref_cv, rel=1e-6)
    f = mesh.forces(d_pt, N, box, torch.tensor([bias], dtype=torch.float64, device="cuda")).cpu().numpy()
    fr = G[name + "_f64_force"]
    # the reference's double build rounds |x| to float in assignTSCderiv (copysignf, OrderParameterMesh.cc:473), which
    # perturbs ITS forces at the 1e-4 level of max|F| on these inputs; the 1e-5 tolerance is checked against the oracle with
    # |x| exact, which equals the reference to 1e-12 with the quirk on (tests/test_reference_build.py)
    assert np.abs(f - fr).max() < 2e-4 * np.abs(fr).max()
    m = oracle.Mesh(*dims, modes, L, N, "f64", literal_copysignf=False)
    m.current_value(pt)
    fo = m.forces(pt, bias)
    assert np.abs(f - fo).max() < 1e-5 * np.abs(fo).max()


@pytest.mark.parametrize("name", ["m0", "m1", "m2"])
def test_mesh_qmax_and_virial_against_reference_vectors(gpu, oracle, name):
    """The epilogues of the fused z sweep (knob 13) against the REFERENCE's own computeQmax and computeVirial
    (tests/golden/ref_golden.npz): q*_max / sq_max (arg-max of |f_k|^2 over all k INCLUDING k = 0, as the reference scans) and
    the k-space virial with a tabulated kernel derivative."""
    import torch
    G = _ref_gold()
    c = G[name + "_cfg"]
    dims, L, bias, modes = tuple(int(v) for v in c[:3]), tuple(c[3:6]), float(c[6]), tuple(c[7:])
    pt = G[name + "_postype"]
    N = pt.shape[0]
    kmin, kmax, n = G["virial_table"]
    kt = np.linspace(kmin, kmax, int(n))
    dK = -2.0 * (kt - 2.0) * np.exp(-(kt - 2.0) ** 2)
    d_pt = torch.from_numpy(pt).cuda()
    box = gpu.Box.make(L)
    mesh = gpu.Mesh(*dims, modes)
    mesh.set(13, 1)
    mesh.set_table(dK, kmin, kmax)
    cv = mesh.compute_cv(d_pt, N, box).cpu().item()
    assert cv == pytest.approx(G[name + "_f64_cv"][0], rel=1e-6)          # the epilogues do not disturb the CV
    x = mesh.extras()
    ref_q = G[name + "_f64_qmax"]
    assert x["sq_max"] == pytest.approx(ref_q[3], rel=2e-6)
    q = x["q_max"]
    assert np.allclose(q, ref_q[:3], rtol=1e-6, atol=1e-12) or np.allclose(q, -ref_q[:3], rtol=1e-6, atol=1e-12)     # k and -k tie by symmetry
    ref_v = G[name + "_f64_virial"]
    np.testing.assert_allclose(bias * x["virial"], ref_v, rtol=2e-5, atol=2e-6 * np.abs(ref_v).max())
    # without the table the reference's virial is identically zero
    mesh.set_table([], 0, 0, use_table=False)
    mesh.compute_cv(d_pt, N, box)
    assert np.all(mesh.extras()["virial"] == 0.0)
    # and with the epilogues off the getter refuses instead of returning stale numbers
    mesh.set(13, 0)
    mesh.compute_cv(d_pt, N, box)
    from metadynamics_plugin_b200._abi import MetadError
    with pytest.raises(MetadError):
        mesh.extras()


def test_mesh_qmax_at_c3_size(gpu, oracle):
    """q_max / sq_max on 128^3 with a lamellar-ordered two-type melt (the peak is a real structure-factor peak, not k = 0)."""
    import torch
    from metadynamics_plugin_b200 import workloads
    N, L = 1 << 18, 64.0
    pos, types = workloads.diblock(N, L, 5, 11)
    pt = gpu.make_postype(pos, types)
    h_pt = host_pt(oracle, pos.astype(np.float32), types)
    mesh = gpu.Mesh(128, 128, 128, [1.0, -1.0])
    mesh.set(13, 1)
    mesh.compute_cv(pt, N, gpu.Box.make(L))
    x = mesh.extras()
    m = oracle.Mesh(128, 128, 128, [1.0, -1.0], L, N, "f64")
    m.current_value(h_pt)
    qo = m.qmax()
    assert x["sq_max"] == pytest.approx(qo[3], rel=2e-6)
    assert np.allclose(np.abs(x["q_max"]), np.abs(qo[:3]), rtol=1e-6, atol=1e-12)
    assert abs(abs(x["q_max"][2]) - 2 * np.pi * 5 / L) < 1e-6 and abs(x["q_max"][0]) < 1e-12       # the imposed lamellar period


@pytest.mark.parametrize("name", ["l0", "l1"])
def test_lamellar_against_reference_vectors(gpu, name):
    """The device path against outputs of the REFERENCE's own LamellarOrderParameter.cc (double build)."""
    import torch
    G = _ref_gold()
    c = G[name + "_cfg"]
    L, tilt, bias, nw = tuple(c[:3]), tuple(c[3:6]), float(c[6]), int(c[7])
    lv = c[8:8 + 3 * nw].astype(int).reshape(-1, 3)
    modes = tuple(c[8 + 3 * nw:])
    pt = G[name + "_postype"]
    N = pt.shape[0]
    d_pt = torch.from_numpy(pt).cuda()
    box = gpu.Box.make(L, tilt)
    lam = gpu.Lamellar(modes, lv)
    cv = lam.compute_modes(d_pt, N, box).cpu().item()
    scale = np.sqrt(N) * max(abs(v) for v in modes)
    assert np.abs(lam.modes.cpu().numpy().reshape(-1, 2) - G[name + "_f64_modes"]).max() < 2e-6 * scale
    assert abs(cv - G[name + "_f64_cv"][0]) < 2e-6 * scale * nw / N
    f = lam.forces(d_pt, N, box, torch.tensor([bias], dtype=torch.float64, device="cuda")).cpu().numpy()
    fr = G[name + "_f64_force"]
    assert np.abs(f - fr).max() < 1e-5 * np.abs(fr).max()


def test_mesh_order_independence(gpu):
    """Shuffling the particle array permutes the forces and leaves density / CV / forces BITWISE unchanged
    (the density is accumulated in integers)."""
    import torch
    N, dims, L = 40000, (64, 64, 64), 30.0
    pos, types = rand_pt(N, L, 2, 5)
    box = gpu.Box.make(L)
    mesh = gpu.Mesh(*dims, [1.0, -0.4])
    mesh.set(1, 1)
    bias = torch.tensor([1.0], dtype=torch.float64, device="cuda")
    d1 = to_dev(gpu, pos, types)
    cv = mesh.compute_cv(d1, N, box).cpu().item()
    rho = mesh.rho()
    f = mesh.forces(d1, N, box, bias).cpu().numpy()
    perm = np.random.default_rng(0).permutation(N)
    d2 = to_dev(gpu, pos[perm], types[perm])
    mesh2 = gpu.Mesh(*dims, [1.0, -0.4])
    mesh2.set(1, 1)
    cv2 = mesh2.compute_cv(d2, N, box).cpu().item()
    f2 = mesh2.forces(d2, N, box, bias).cpu().numpy()
    assert np.array_equal(mesh2.rho(), rho)
    assert cv2 == cv
    assert np.array_equal(f2, f[perm])


def test_mesh_full_size_c3(gpu, oracle):
    """BASELINE configuration C3 at its full size (N = 2^20 particles, 128^3 mesh, the bench's own workload generator):
    direct parity with the double oracle (which needs a few seconds at this size), bit-exact cell indices, and the
    size-independent properties of the path: mass conservation of the assignment (TSC weights sum to one), exact linearity
    of the force in the bias factor, bitwise independence of the particle order."""
    import torch
    from metadynamics_plugin_b200 import workloads
    w = workloads.c3()
    pt = w["postype"]
    N, dims, L, modes = pt.shape[0], w["mesh"], w["L"], w["mode"]
    assert N == 1 << 20 and tuple(dims) == (128, 128, 128)
    box = gpu.Box.make(L)
    d_pt = torch.from_numpy(pt).cuda()
    mesh = gpu.Mesh(*dims, modes)
    mesh.set(1, 1)
    mesh.set(3, 1)
    cv = mesh.compute_cv(d_pt, N, box).cpu().item()
    one = torch.tensor([1.0], dtype=torch.float64, device="cuda")
    f = mesh.forces(d_pt, N, box, one).cpu().numpy()
    # parity at full size
    m = oracle.Mesh(*dims, modes, L, N, "f64", literal_copysignf=False)
    cvo = m.current_value(pt)
    fo = m.forces(pt, 1.0)
    m32 = oracle.Mesh(*dims, modes, L, N, "f32")
    m32.assign(pt)
    assert np.array_equal(mesh.cells(), m32.cells())
    assert cv == pytest.approx(cvo, rel=1e-6)
    assert np.abs(f - fo).max() < 1e-5 * np.abs(fo).max()
    # mass conservation
    assert abs(np.asarray(mesh.rho(), dtype=np.float64).sum() - N) < 1e-6 * N
    # linearity in the bias factor: doubling it doubles every force component exactly
    mesh.compute_cv(d_pt, N, box)
    f2 = mesh.forces(d_pt, N, box, 2.0 * one).cpu().numpy()
    assert np.array_equal(f2, 2.0 * f)
    # particle order
    perm = np.random.default_rng(3).permutation(N)
    d_pp = torch.from_numpy(np.ascontiguousarray(pt[perm])).cuda()
    other = gpu.Mesh(*dims, modes)
    cvp = other.compute_cv(d_pp, N, box).cpu().item()
    fp = other.forces(d_pp, N, box, one).cpu().numpy()
    assert cvp == cv
    assert np.array_equal(fp, f[perm])


# every FFT line-length instantiation the plan can dispatch (x: nx/2 in 16..512, y and z: 16..512) on the DEVICE, against the
# double oracle -- the CPU emulation (tests/test_emulation.py) checks the index logic of the same lengths, not the compiled code
FFT_LEN_CASES = [
    (32, 256, 16), (32, 16, 256), (512, 16, 16), (1024, 16, 16), (32, 512, 16), (32, 16, 512),
    (64, 128, 32), (256, 32, 128), (128, 64, 256),
    (128, 128, 32), (256, 256, 16), (512, 512, 16),          # planes with a fused x+y instantiation (clusters of 1, 2, 8 CTAs)
]


@pytest.mark.parametrize("dims", FFT_LEN_CASES)
def test_mesh_every_fft_length(gpu, oracle, dims):
    import torch
    N = 30000
    Lf = np.asarray(dims, float) * 0.31
    pos, types = rand_pt(N, Lf, 2, sum(dims))
    modes = (1.0, -0.7)
    d_pt = to_dev(gpu, pos, types)
    h_pt = host_pt(oracle, pos, types)
    box = gpu.Box.make(Lf)
    mesh = gpu.Mesh(*dims, modes)
    mesh.set(3, 1)
    cv = mesh.compute_cv(d_pt, N, box).cpu().item()
    m = oracle.Mesh(*dims, modes, Lf, N, "f64", literal_copysignf=False)
    cvo = m.current_value(h_pt)
    assert cv == pytest.approx(cvo, rel=1e-6)
    # Re IFFT(G) up to its mean (the device removes the k = 0 mode, which cannot produce a force)
    inv, inv_o = np.asarray(mesh.inv(), dtype=np.float64), m.inv_re
    inv -= inv.mean(); inv_o = inv_o - inv_o.mean()
    assert np.abs(inv - inv_o).max() < 2e-5 * np.abs(inv_o).max()
    bias = torch.tensor([0.77], dtype=torch.float64, device="cuda")
    f = mesh.forces(d_pt, N, box, bias).cpu().numpy()
    fo = m.forces(h_pt, 0.77)
    assert np.abs(f - fo).max() < 1e-5 * np.abs(fo).max()
    m32 = oracle.Mesh(*dims, modes, Lf, N, "f32")
    m32.assign(h_pt)
    assert np.array_equal(mesh.cells(), m32.cells())


@pytest.mark.parametrize("dims", [(128, 128, 32), (256, 256, 32), (512, 512, 16)])
def test_mesh_fused_xy_equals_separate_sweeps(gpu, dims):
    """Knob 15: the x and y sweeps fused on whole z planes by thread-block clusters (distributed shared memory) perform the same
    per-line arithmetic as the separate sweeps: CV, Re IFFT(G) and forces agree to rounding (here: bit for bit or a few ulps)."""
    import torch
    N = 200000
    Lf = np.asarray(dims, float) * 0.9
    pos, types = rand_pt(N, Lf, 2, 31)
    d_pt = to_dev(gpu, pos, types)
    box = gpu.Box.make(Lf)
    bias = torch.tensor([0.9], dtype=torch.float64, device="cuda")
    res = []
    for fused in (2, 0):                                  # 2: every plane shape with a fused instantiation; 0: separate sweeps
        mesh = gpu.Mesh(*dims, (1.0, -0.6))
        mesh.set(15, fused)
        cv = mesh.compute_cv(d_pt, N, box).cpu().item()
        inv = np.asarray(mesh.inv(), dtype=np.float64)
        f = mesh.forces(d_pt, N, box, bias).cpu().numpy()
        cv2 = mesh.compute_cv(d_pt, N, box).cpu().item()          # the accumulator was cleared by the fused forward sweep
        assert cv2 == cv
        res.append((cv, inv, f))
    assert res[0][0] == pytest.approx(res[1][0], rel=1e-9)
    assert np.abs(res[0][1] - res[1][1]).max() <= 1e-6 * np.abs(res[1][1]).max()
    assert np.abs(res[0][2] - res[1][2]).max() <= 1e-6 * np.abs(res[1][2]).max()


def test_mesh_full_size_c4(gpu, oracle):
    """BASELINE configuration C4 -- the headline -- at its full size (N = 2^24, 256^3 mesh, the bench's own generator):
    direct parity with the double oracle (CV 1e-6, forces 1e-5 of max|F|), cell indices bit-exact against the float
    oracle, mass conservation."""
    import torch
    from metadynamics_plugin_b200 import workloads
    w = workloads.c4()
    pt = w["postype"]
    N, dims, L, modes = pt.shape[0], w["mesh"], w["L"], w["mode"]
    assert N == 1 << 24 and tuple(dims) == (256, 256, 256)
    box = gpu.Box.make(L)
    d_pt = torch.from_numpy(pt).cuda()
    mesh = gpu.Mesh(*dims, modes)
    mesh.set(1, 1)
    mesh.set(3, 1)
    cv = mesh.compute_cv(d_pt, N, box).cpu().item()
    one = torch.tensor([1.0], dtype=torch.float64, device="cuda")
    f = mesh.forces(d_pt, N, box, one).cpu().numpy()
    cells = mesh.cells()
    rho_sum = np.asarray(mesh.rho(), dtype=np.float64).sum()
    del mesh, d_pt
    m32 = oracle.Mesh(*dims, modes, L, N, "f32")
    m32.assign(pt)
    assert np.array_equal(cells, m32.cells())
    del m32, cells
    m = oracle.Mesh(*dims, modes, L, N, "f64", literal_copysignf=False)
    cvo = m.current_value(pt)
    fo = m.forces(pt, 1.0)
    assert cv == pytest.approx(cvo, rel=1e-6)
    assert np.abs(f - fo).max() < 1e-5 * np.abs(fo).max()
    assert abs(rho_sum - N) < 1e-6 * N


def test_lamellar_full_size_c2_c5(gpu, oracle):
    """Configurations C2 (N = 262 144) and C5 (N = 2^23) at full size: CV 1e-6 relative, forces 1e-5 of max|F|."""
    import torch
    from metadynamics_plugin_b200 import workloads
    for w in (workloads.c2(), workloads.c5()):
        pt = w["postype"]
        N = pt.shape[0]
        d_pt = torch.from_numpy(pt).cuda()
        box = gpu.Box.make(w["L"])
        lam = gpu.Lamellar(w["mode"], w["lattice_vectors"])
        cv = lam.compute_modes(d_pt, N, box).cpu().item()
        cvo, _ = oracle.lamellar_cv(pt, N, w["mode"], w["lattice_vectors"], w["L"])
        assert abs(cvo) > 0.05
        assert cv == pytest.approx(cvo, rel=1e-6), w["name"]
        bias = torch.tensor([-0.37], dtype=torch.float64, device="cuda")
        f = lam.forces(d_pt, N, box, bias).cpu().numpy()
        fo = oracle.lamellar_forces(pt, N, w["mode"], w["lattice_vectors"], w["L"], -0.37)
        assert np.abs(f - fo).max() < 1e-5 * np.abs(fo).max(), w["name"]


@pytest.mark.parametrize("dims,N", [((64, 64, 64), 70000), ((128, 128, 128), 300000)])
def test_mesh_bank_order_equals_layer_order(gpu, dims, N):
    """The two orders of the particles inside a tile (knob 6: bank order, layer order) are permutations of one another:
    density, CV and forces are bitwise equal (8-cell and 16-cell tiles)."""
    import torch
    L = 30.0
    pos, types = rand_pt(N, L, 2, 5)
    modes = [1.0, -0.5]
    box = gpu.Box.make(L)
    bias = torch.tensor([1.1], dtype=torch.float64, device="cuda")
    d_pt = to_dev(gpu, pos, types)
    out = []
    for kind in (1, 0):
        mesh = gpu.Mesh(*dims, modes)
        mesh.set(6, kind)
        mesh.set(1, 1)
        cv = mesh.compute_cv(d_pt, N, box).cpu().item()
        rho = np.asarray(mesh.rho())
        f = mesh.forces(d_pt, N, box, bias).cpu().numpy()
        assert mesh.stats()["drifted"] == 0
        out.append((cv, f, rho))
    assert out[0][0] == out[1][0]
    assert np.array_equal(out[0][1], out[1][1])
    assert np.array_equal(out[0][2], out[1][2])


def test_mesh_stale_tile_order(gpu, oracle):
    """The tile order is reused across calls while the particles move: every call is still exact (cells are recomputed
    from the current positions), drifted particles take the direct path and trigger a rebuild; results are bitwise those
    of a fresh plan."""
    import torch
    N, dims, L = 60000, (64, 64, 64), 40.0
    rng = np.random.default_rng(11)
    pos, types = rand_pt(N, L, 2, 3)
    modes = [1.0, -1.0]
    box = gpu.Box.make(L)
    bias = torch.tensor([0.9], dtype=torch.float64, device="cuda")
    mesh = gpu.Mesh(*dims, modes)
    mesh.set(0, 1000)                                    # never rebuild on the period
    mesh.set(3, 1)
    h = L / dims[0]
    rebuilds = []
    for step, amp in enumerate((0.0, 0.3, 0.3, 0.3, 2.5, 0.0, 0.0)):      # displacement per step in cells
        pos = pos + (rng.random((N, 3)).astype(np.float32) - 0.5) * np.float32(2 * amp * h)
        pos = (((pos + L / 2) % L) - L / 2).astype(np.float32)
        pos[pos >= np.float32(L / 2)] = -np.float32(L / 2)
        d_pt = to_dev(gpu, pos, types)
        cv = mesh.compute_cv(d_pt, N, box).cpu().item()
        f = mesh.forces(d_pt, N, box, bias).cpu().numpy()
        st = mesh.stats()
        rebuilds.append(st["rebuilds"])
        fresh = gpu.Mesh(*dims, modes)
        cvf = fresh.compute_cv(d_pt, N, box).cpu().item()
        ff = fresh.forces(d_pt, N, box, bias).cpu().numpy()
        assert cv == cvf and np.array_equal(f, ff), step
        if step in (0, 4):
            h_pt = host_pt(oracle, pos, types)
            m = oracle.Mesh(*dims, modes, [L] * 3, N, "f64", literal_copysignf=False)
            assert cv == pytest.approx(m.current_value(h_pt), rel=1e-6)
            fo = m.forces(h_pt, 0.9)
            assert np.abs(f - fo).max() < 1e-5 * np.abs(fo).max()
            m32 = oracle.Mesh(*dims, modes, [L] * 3, N, "f32")
            m32.assign(h_pt)
            assert np.array_equal(mesh.cells(), m32.cells())
        if step == 4:
            assert st["drifted"] > N // 256              # the big move left many particles outside their padded tile
    assert rebuilds[0] == 1 and rebuilds[4] == 1         # small moves reuse the order ...
    assert rebuilds[-1] == 2                             # ... the drift report triggers exactly one rebuild
    mesh.set(0, 2)                                       # periodic rebuild
    for _ in range(4):
        mesh.compute_cv(d_pt, N, box)
    assert mesh.stats()["rebuilds"] == 4


def test_mesh_cuda_graph_replay(gpu):
    """Knob 4: the kernel sequence of metad_mesh_cv is captured into a CUDA graph and replayed.  Positions change in place
    between the calls, the tile order is rebuilt on its period (outside the graph): every call must be bitwise equal to a
    fresh eager plan."""
    import torch
    N, dims, L = 50000, (64, 64, 64), 35.0
    rng = np.random.default_rng(21)
    pos, types = rand_pt(N, L, 2, 9)
    modes = [1.0, -0.5]
    box = gpu.Box.make(L)
    bias = torch.tensor([1.1], dtype=torch.float64, device="cuda")
    mesh = gpu.Mesh(*dims, modes)
    mesh.set(4, 1)
    mesh.set(0, 3)                                       # rebuild every third call
    d_pt = to_dev(gpu, pos, types)
    for step in range(8):
        pos = pos + (rng.random((N, 3)).astype(np.float32) - 0.5) * np.float32(0.2)
        pos = (((pos + L / 2) % L) - L / 2).astype(np.float32)
        pos[pos >= np.float32(L / 2)] = -np.float32(L / 2)
        d_pt.copy_(to_dev(gpu, pos, types))              # same device buffer, new contents
        cv = mesh.compute_cv(d_pt, N, box).cpu().item()
        f = mesh.forces(d_pt, N, box, bias).cpu().numpy()
        fresh = gpu.Mesh(*dims, modes)
        cvf = fresh.compute_cv(d_pt, N, box).cpu().item()
        ff = fresh.forces(d_pt, N, box, bias).cpu().numpy()
        assert cv == cvf and np.array_equal(f, ff), step
    assert mesh.graph_launches() >= 6                   # call 0 eager, call 1 captures + replays, then replays
    assert mesh.stats()["rebuilds"] == 3
    # a different buffer is a different signature: falls back to eager, then a new graph
    d2 = d_pt.clone()
    assert mesh.compute_cv(d2, N, box).cpu().item() == cv


@pytest.mark.parametrize("sigma,N", [(3.0, 200000), (0.8, 200000), (0.35, 1500000)])
def test_mesh_dense_cells_wide_accumulation(gpu, oracle, sigma, N):
    """Many particles per cell and large mode coefficients.  A 32-bit fixed-point density shares its range between the
    resolution of one tap and the total of a cell; once the largest cell load would cost resolution the plan switches to
    64-bit accumulation (split 32-bit tiles in shared memory, 64-bit mesh), so the north star's 1e-6 on the CV holds at any
    density: sigma = 3 -> ~25 particles in the densest cell, 0.8 -> ~800, 0.35 with N = 1.5 M -> ~20 000."""
    import torch
    dims, L = (32, 32, 32), 8.0
    rng = np.random.default_rng(5)
    pos = (rng.normal(0.0, sigma, (N, 3))).astype(np.float32)
    pos = (((pos + L / 2) % L) - L / 2).astype(np.float32)
    pos[pos >= np.float32(L / 2)] = -np.float32(L / 2)
    types = rng.integers(0, 2, N).astype(np.int32)
    modes = [250.0, 100.0]
    box = gpu.Box.make(L)
    mesh = gpu.Mesh(*dims, modes)
    mesh.set(1, 1)
    mesh.set(3, 1)
    d_pt = to_dev(gpu, pos, types)
    cv = mesh.compute_cv(d_pt, N, box).cpu().item()
    st, acc = mesh.stats(), mesh.accumulator()
    h_pt = host_pt(oracle, pos, types)
    m32 = oracle.Mesh(*dims, modes, [L] * 3, N, "f32")
    m32.assign(h_pt)
    c = m32.cells()
    assert np.array_equal(mesh.cells(), c)
    max_count = np.bincount(c[:, 0] + 32 * (c[:, 1] + 32 * c[:, 2])).max()
    assert st["range_warnings"] == 0 and max_count >= 13
    assert acc["wide"] and acc["requested"] == 2
    # the scale is set by the largest tap alone: resolution 2^-23 of one particle's contribution at any density
    if acc["wide"]:
        assert st["fx_scale"] == 2.0 ** np.floor(np.log2(2 ** 22 / (0.421875 * 250.0)))
    m = oracle.Mesh(*dims, modes, [L] * 3, N, "f64", literal_copysignf=False)
    cvo = m.current_value(h_pt)
    assert np.abs(mesh.rho() - m.mesh).max() < 1e-6 * np.abs(m.mesh).max()
    assert cv == pytest.approx(cvo, rel=1e-6)
    bias = torch.tensor([1.0], dtype=torch.float64, device="cuda")
    f = mesh.forces(d_pt, N, box, bias).cpu().numpy()
    fo = m.forces(h_pt, 1.0)
    assert np.abs(f - fo).max() < 1e-5 * np.abs(fo).max()
    # the width is re-decided at every rebuild of the tile order: a dilute set of particles goes back to 32 bits
    pos2, types2 = rand_pt(20000, L, 2, 8)
    mesh.compute_cv(to_dev(gpu, pos2, types2), 20000, box)
    assert not mesh.accumulator()["wide"]


def test_mesh_accumulator_follows_density_without_sync(gpu, oracle):
    """Same particle number, positions change from dilute to clustered in place: the drift report triggers a rebuild, the
    rebuild reports (asynchronously) that 64-bit accumulation is needed, the next call switches.  Every call in between is
    still within the tolerance that a 32-bit density allows; from the switch on the CV is back at 1e-6."""
    dims, L, N = (32, 32, 32), 8.0, 60000
    box = gpu.Box.make(L)
    modes = [1.0, -1.0]
    pos, types = rand_pt(N, L, 2, 4)
    mesh = gpu.Mesh(*dims, modes)
    mesh.set(0, 1000)
    d_pt = to_dev(gpu, pos, types)
    mesh.compute_cv(d_pt, N, box)
    assert not mesh.accumulator()["wide"]
    rng = np.random.default_rng(9)
    posc = rng.normal(0.0, 0.5, (N, 3)).astype(np.float32)
    posc = (((posc + L / 2) % L) - L / 2).astype(np.float32)
    d_pt.copy_(to_dev(gpu, posc, types))
    h_pt = host_pt(oracle, posc, types)
    cvo = oracle.Mesh(*dims, modes, [L] * 3, N, "f64", literal_copysignf=False).current_value(h_pt)
    cvs = []
    for _ in range(5):
        cvs.append(mesh.compute_cv(d_pt, N, box).cpu().item())      # .item() synchronises the TEST, not the library
    assert mesh.accumulator()["wide"]
    assert cvs[-1] == pytest.approx(cvo, rel=1e-6)
    assert all(c == pytest.approx(cvo, rel=1e-4) for c in cvs)


def test_mesh_c1_golden_with_umbrella(gpu, oracle):
    """Config C1 (the reference's test/test_mesh.py geometry): CV, device-side harmonic umbrella, forces."""
    import json, os, torch
    from metadynamics_plugin_b200 import workloads
    here = os.path.dirname(os.path.abspath(__file__))
    gold = json.load(open(os.path.join(here, "golden", "golden.json")))
    data = np.load(os.path.join(here, "golden", "golden.npz"))
    w = workloads.c1()
    d_pt = torch.from_numpy(w["postype"]).cuda()
    box = gpu.Box.make(w["L"])
    mesh = gpu.Mesh(*w["mesh"], w["mode"])
    mesh.set(3, 1)
    cv = mesh.compute_cv(d_pt, 1000, box)
    assert cv.cpu().item() == pytest.approx(gold["c1_cv"], rel=1e-6)
    u = w["umbrella"]
    energy = torch.zeros(1, dtype=torch.float64, device="cuda")
    bias = gpu.umbrella_apply("harmonic", cv, None, cv0=u["cv0"], kappa=u["kappa"], energy_out=energy)
    assert bias.cpu().item() == pytest.approx(gold["c1_bias"], rel=1e-4)       # kappa*(cv-cv0): cancellation amplifies 1e-6
    assert energy.cpu().item() == pytest.approx(gold["c1_umbrella_energy"], rel=2e-4)
    # forces with the oracle's bias (isolates the force kernel from the umbrella's cancellation)
    b = torch.tensor([gold["c1_bias"]], dtype=torch.float64, device="cuda")
    f = mesh.forces(d_pt, 1000, box, b).cpu().numpy()
    assert np.abs(f - data["c1_forces"]).max() < 1e-5 * np.abs(data["c1_forces"]).max()
    assert np.array_equal(mesh.cells(), data["c1_cells"])


def test_mesh_empty_and_tiny_inputs(gpu, oracle):
    """N = 0 (an MPI rank without particles) and N = 1: no kernel may trip over empty tiles."""
    import torch
    box = gpu.Box.make(6.0)
    mesh = gpu.Mesh(32, 32, 32, [1.0])
    empty = torch.empty((0, 4), dtype=torch.float32, device="cuda")
    cv0 = mesh.compute_cv(empty, 5, box).cpu().item()
    assert cv0 == 0.0
    f0 = mesh.forces(empty, 5, box, torch.ones(1, dtype=torch.float64, device="cuda"))
    assert f0.shape == (0, 4)
    pos = np.array([[0.3, -1.2, 2.9]], np.float32)
    d1 = to_dev(gpu, pos, np.zeros(1, np.int32))
    cv1 = mesh.compute_cv(d1, 1, box).cpu().item()
    m = oracle.Mesh(32, 32, 32, [1.0], [6.0] * 3, 1, "f64", literal_copysignf=False)
    assert cv1 == pytest.approx(m.current_value(host_pt(oracle, pos, np.zeros(1, np.int32))), rel=1e-6)
    assert mesh.compute_cv(empty, 5, box).cpu().item() == 0.0        # and back to empty: the accumulator was cleared


# ------------------------------------------------------------------------------------------------ mesh CV in triclinic boxes
def test_mesh_triclinic_against_reference_vector(gpu, oracle):
    """Vector t0 (tests/golden/ref_golden.npz): the REFERENCE's own OrderParameterMesh.cc, double build, in a box with tilt
    factors (0.2, -0.1, 0.15).  Its in-cell offsets carry a constant there (makeFraction(shift + lo) shears `lo` as well,
    OrderParameterMesh.cc:571-573, 806-808: 0.67 cells along x, 1.33 along y) and TSC weight beyond |x| = 3/2 is dropped;
    the device path reproduces that behaviour by default (knob 16)."""
    import torch
    G = _ref_gold()
    c = G["t0_cfg"]
    dims, L, tilt, bias, modes = tuple(int(v) for v in c[:3]), tuple(c[3:6]), tuple(c[6:9]), float(c[9]), tuple(c[10:])
    pt = G["t0_postype"]
    N = pt.shape[0]
    d_pt = torch.from_numpy(pt).cuda()
    box = gpu.Box.make(L, tilt)
    mesh = gpu.Mesh(*dims, modes)
    mesh.set(1, 1)
    mesh.set(3, 1)
    mesh.set(13, 1)                                                         # q_max epilogue (checked against the oracle below)
    cv = mesh.compute_cv(d_pt, N, box).cpu().item()
    ref_cv, ref_msq = G["t0_f64_cv"]
    assert mesh.mode_sq() == ref_msq
    rho = G["t0_f64_rho"]
    assert np.abs(rho.sum()) < 0.9 * N                                      # the reference does lose weight here
    assert np.abs(mesh.rho() - rho).max() < 2e-6 * max(1.0, np.abs(rho).max())
    assert cv == pytest.approx(ref_cv, rel=1e-6)
    f = mesh.forces(d_pt, N, box, torch.tensor([bias], dtype=torch.float64, device="cuda")).cpu().numpy()
    fr = G["t0_f64_force"]
    assert np.abs(f - fr).max() < 2e-4 * np.abs(fr).max()                   # copysignf in the reference's double build, see above
    m = oracle.Mesh(*dims, modes, L, N, "f64", tilt=tilt, literal_copysignf=False)
    m.current_value(pt)
    fo = m.forces(pt, bias)
    assert np.abs(f - fo).max() < 1e-5 * np.abs(fo).max()
    m32 = oracle.Mesh(*dims, modes, L, N, "f32", tilt=tilt)
    m32.assign(pt)
    assert np.array_equal(mesh.cells(), m32.cells())                        # single-precision BoxDim::makeFraction, bit for bit
    # computeQmax scans k = 0 as well, and with weight lost f_0 is no longer sum a / N: the kernel restores what the mean
    # removal took out (ConvParams::dc_restore)
    x, qo = mesh.extras(), m.qmax()
    assert x["sq_max"] == pytest.approx(qo[3], rel=2e-6)
    assert np.allclose(x["q_max"], qo[:3], rtol=1e-6, atol=1e-12) or np.allclose(x["q_max"], -qo[:3], rtol=1e-6, atol=1e-12)


TRI_CASES = [
    (3000, (32, 32, 32), (10.0, 10.0, 10.0), (0.2, -0.1, 0.15), (1.0,)),
    (5000, (32, 16, 64), (10.0, 7.3, 21.1), (-0.35, 0.4, 0.25), (1.0, -1.0)),       # literal offset 12 cells: all weight lost
    (40000, (64, 64, 64), (30.0, 28.0, 33.0), (0.02, -0.01, 0.03), (1.0, -1.0)),    # small tilt: offsets below one cell
    (300000, (128, 64, 128), (70.0, 35.0, 70.0), (0.3, 0.1, -0.2), (1.0, -0.5)),    # 16^3 tiles
]


@pytest.mark.parametrize("literal", [True, False])
@pytest.mark.parametrize("N,dims,L,tilt,modes", TRI_CASES)
def test_mesh_triclinic_cv_forces_cells(gpu, oracle, N, dims, L, tilt, modes, literal):
    """cv.mesh in a triclinic box against the oracle: with the reference's literal in-cell offsets (default) and with the
    geometrically correct ones (knob 16 = 0; that assignment is continuous, so particles within ulps of cell faces are part
    of the input)."""
    import torch
    from conftest import triclinic_case
    pos, types = triclinic_case(N, L, tilt, len(modes), N % 1000 + 7, faces=not literal)
    d_pt = to_dev(gpu, pos, types)
    h_pt = host_pt(oracle, pos, types)
    box = gpu.Box.make(L, tilt)
    mesh = gpu.Mesh(*dims, modes)
    mesh.set(16, 1 if literal else 0)
    mesh.set(1, 1)
    mesh.set(3, 1)
    cv = mesh.compute_cv(d_pt, N, box).cpu().item()
    m = oracle.Mesh(*dims, modes, L, N, "f64", tilt=tilt, literal_copysignf=False, literal_tilt_offset=literal)
    cvo = m.current_value(h_pt)
    m32 = oracle.Mesh(*dims, modes, L, N, "f32", tilt=tilt)
    m32.assign(h_pt)
    assert np.array_equal(mesh.cells(), m32.cells())
    assert mesh.mode_sq() == m.mode_sq()
    assert np.abs(mesh.rho() - m.mesh).max() < 2e-6 * max(1.0, np.abs(m.mesh).max())
    # literal offsets of more than a cell drop most of the weight; the fixed-point resolution of a tap is then a larger
    # fraction of the density and the CV goes with its fourth power (tests/test_emulation.py)
    lost = literal and abs(m.mesh.sum()) < 0.5 * abs(np.asarray(modes)[types].sum())
    assert cv == pytest.approx(cvo, rel=3e-6 if lost else 1e-6)
    bias = torch.tensor([-0.6], dtype=torch.float64, device="cuda")
    f = mesh.forces(d_pt, N, box, bias).cpu().numpy()
    fo = m.forces(h_pt, -0.6)
    assert np.abs(f - fo).max() <= 1e-5 * np.abs(fo).max()
    assert np.all(f[:, 3] == 0)
    cv2 = mesh.compute_cv(d_pt, N, box).cpu().item()
    f2 = mesh.forces(d_pt, N, box, bias).cpu().numpy()
    assert cv2 == cv and np.array_equal(f, f2)
    st = mesh.stats()
    assert st["rebuilds"] == 1 and st["drifted"] == 0 and st["outside_slab"] == 0
    # without the particle cache (knob 9 = 0) the gather recomputes cell and offsets: same forces
    nocache = gpu.Mesh(*dims, modes)
    nocache.set(16, 1 if literal else 0)
    nocache.set(9, 0)
    assert nocache.compute_cv(d_pt, N, box).cpu().item() == pytest.approx(cv, rel=1e-7)
    assert np.abs(nocache.forces(d_pt, N, box, bias).cpu().numpy() - f).max() <= 1e-6 * np.abs(f).max()


def test_mesh_triclinic_stale_order_and_epilogues(gpu, oracle):
    """Triclinic box: particles that drifted out of their padded tile (direct path of spread and gather), and the q_max /
    virial epilogues with the reciprocal lattice vectors of the sheared box."""
    import torch
    from conftest import triclinic_case
    N, dims, L, tilt, modes = 50000, (64, 64, 64), (40.0, 36.0, 44.0), (0.25, -0.15, 0.1), (1.0, -1.0)
    pos, types = triclinic_case(N, L, tilt, 2, 5, faces=False)
    box = gpu.Box.make(L, tilt)
    bias = torch.tensor([0.9], dtype=torch.float64, device="cuda")
    mesh = gpu.Mesh(*dims, modes)
    mesh.set(16, 0)
    mesh.set(0, 1000)
    mesh.set(13, 1)
    kt = np.linspace(0.2, 6.0, 64)
    dK = -2.0 * (kt - 2.0) * np.exp(-(kt - 2.0) ** 2)
    mesh.set_table(dK, 0.2, 6.0)
    mesh.compute_cv(to_dev(gpu, pos, types), N, box)
    # move every particle by up to 2.5 cells along z (a lattice direction that needs no re-wrapping in x and y for the
    # particles that stay inside) and put it back into the box the way BoxDim::wrap does
    rng = np.random.default_rng(3)
    Lz, (xy, xz, yz) = L[2], tilt
    p = pos.astype(np.float64)
    dz = (rng.random(N) - 0.5) * 5.0 * Lz / dims[2]          # along the lattice vector a3 = (xz, yz, 1) Lz: only the z fraction changes
    p[:, 2] += dz; p[:, 1] += yz * dz; p[:, 0] += xz * dz
    up, dn = p[:, 2] >= Lz / 2, p[:, 2] < -Lz / 2
    for sel, sgn in ((up, -1.0), (dn, 1.0)):
        p[sel, 2] += sgn * Lz; p[sel, 1] += sgn * Lz * yz; p[sel, 0] += sgn * Lz * xz
    pos2 = p.astype(np.float32)
    # keep what is still strictly inside the box after the rounding to float (HOOMD would wrap the rest)
    q = pos2.astype(np.float64)
    fy = (q[:, 1] - yz * q[:, 2]) / L[1] + 0.5
    fx = (q[:, 0] - (xz - yz * xy) * q[:, 2] - xy * q[:, 1]) / L[0] + 0.5
    fz = q[:, 2] / Lz + 0.5
    keep = np.all([(v > 1e-6) & (v < 1.0 - 1e-6) for v in (fx, fy, fz)], axis=0)
    assert keep.sum() > 0.99 * N
    pos2, types2 = pos2[keep], types[keep]
    N2 = pos2.shape[0]
    mesh2 = gpu.Mesh(*dims, modes)
    mesh2.set(16, 0); mesh2.set(0, 1000); mesh2.set(13, 1); mesh2.set_table(dK, 0.2, 6.0)
    d0 = to_dev(gpu, pos[keep], types2)
    mesh2.compute_cv(d0, N2, box)                                            # tile order from the old positions
    d_pt = to_dev(gpu, pos2, types2)
    cv = mesh2.compute_cv(d_pt, N2, box).cpu().item()
    f = mesh2.forces(d_pt, N2, box, bias).cpu().numpy()
    st = mesh2.stats()
    assert st["rebuilds"] == 1 and st["drifted"] > N2 // 256
    h_pt = host_pt(oracle, pos2, types2)
    m = oracle.Mesh(*dims, modes, L, N2, "f64", tilt=tilt, literal_copysignf=False, literal_tilt_offset=False)
    assert cv == pytest.approx(m.current_value(h_pt), rel=1e-6)
    fo = m.forces(h_pt, 0.9)
    assert np.abs(f - fo).max() < 1e-5 * np.abs(fo).max()
    x = mesh2.extras()
    qo = m.qmax()
    assert x["sq_max"] == pytest.approx(qo[3], rel=2e-6)
    assert np.allclose(x["q_max"], qo[:3], rtol=1e-6, atol=1e-12) or np.allclose(x["q_max"], -qo[:3], rtol=1e-6, atol=1e-12)
    vo = m.virial(dK, 0.2, 6.0, 0.9)
    # an ideal gas: the off-diagonal sums cancel to a tenth of the diagonal ones, so the absolute scale is the largest component
    np.testing.assert_allclose(0.9 * x["virial"], vo, rtol=2e-5, atol=1e-5 * np.abs(vo).max())


# ------------------------------------------------------------------------------------------------ mesh CV on any mesh size
GENERAL_CASES = [
    (2000, (24, 20, 18), (11.0, 9.5, 8.0), (1.0,), True),              # 4.2.3 | 4.5 | 2.3.3
    (30000, (48, 30, 36), (10.0, 7.3, 21.1), (1.0, -1.0), True),
    (1500, (7, 11, 13), (6.0, 7.0, 8.0), (1.0, -0.5, 2.0), True),      # prime lengths: one radix-n stage
    (500, (3, 1, 2), (4.0, 5.0, 6.0), (1.0,), False),                  # taps alias onto the same cells
    (200000, (96, 80, 100), (50.0, 40.0, 52.0), (1.0, -1.0), False),
    (5000, (16, 16, 16), (8.0, 8.0, 8.0), (1.0,), True),               # a power of two below the tiled path's range
    (1000, (1021, 2, 3), (100.0, 3.0, 4.0), (1.0,), False),            # the longest prime line
]


@pytest.mark.parametrize("N,dims,L,modes,edge", GENERAL_CASES)
def test_mesh_general_path_any_mesh_size(gpu, oracle, N, dims, L, modes, edge):
    """Mesh sizes that are not powers of two (or outside the tiled kernels' range) take the general path
    (csrc/mesh_general.cuh); the reference accepts any size on one rank (OrderParameterMesh.cc:70-79)."""
    import torch
    Lf = np.asarray(L, float)
    pos, types = rand_pt(N, Lf, len(modes), N % 1000 + 3)
    if edge:
        pos[0] = [np.float32(Lf[0]) / 2, 0, 0]
        pos[1] = [-np.float32(Lf[0]) / 2, np.float32(Lf[1]) / 2, -np.float32(Lf[2]) / 2]
        pos[2] = np.nextafter((Lf / 2).astype(np.float32), np.float32(0))
    d_pt = to_dev(gpu, pos, types)
    h_pt = host_pt(oracle, pos, types)
    box = gpu.Box.make(Lf)
    mesh = gpu.Mesh(*dims, modes)
    mesh.set(1, 1)
    mesh.set(3, 1)
    cv = mesh.compute_cv(d_pt, N, box).cpu().item()
    m = oracle.Mesh(*dims, modes, Lf, N, "f64", literal_copysignf=False)
    cvo = m.current_value(h_pt)
    m32 = oracle.Mesh(*dims, modes, Lf, N, "f32")
    m32.assign(h_pt)
    assert np.array_equal(mesh.cells(), m32.cells())
    assert mesh.mode_sq() == m.mode_sq()
    assert np.abs(mesh.rho() - m.mesh).max() < 2e-6 * max(1.0, np.abs(m.mesh).max())
    assert cv == pytest.approx(cvo, rel=1e-6)
    dinv = mesh.inv() - m.inv_re
    dinv -= dinv.mean()
    assert np.abs(dinv).max() < 5e-6 * np.abs(m.inv_re - m.inv_re.mean()).max()
    bias = torch.tensor([0.61], dtype=torch.float64, device="cuda")
    f = mesh.forces(d_pt, N, box, bias).cpu().numpy()
    fo = m.forces(h_pt, 0.61)
    assert np.abs(f - fo).max() <= 1e-5 * np.abs(fo).max()
    assert np.all(f[:, 3] == 0)
    # integer accumulation and fixed summation orders: a second evaluation, and a shuffled input, are bitwise identical
    cv2 = mesh.compute_cv(d_pt, N, box).cpu().item()
    f2 = mesh.forces(d_pt, N, box, bias).cpu().numpy()
    assert cv2 == cv and np.array_equal(f, f2)
    perm = np.random.default_rng(1).permutation(N)
    d_sh = to_dev(gpu, pos[perm], types[perm])
    rho1 = mesh.rho().copy()
    mesh.compute_cv(d_sh, N, box)
    assert np.array_equal(mesh.rho(), rho1)


@pytest.mark.parametrize("name", ["t1", "t2"])
def test_mesh_general_path_against_reference_vectors(gpu, oracle, name):
    """Vectors t1 (24 x 20 x 18, orthorhombic) and t2 (20 x 12 x 18, triclinic) of the REFERENCE's own OrderParameterMesh.cc."""
    import torch
    G = _ref_gold()
    c = G[name + "_cfg"]
    dims, L, tilt, bias, modes = tuple(int(v) for v in c[:3]), tuple(c[3:6]), tuple(c[6:9]), float(c[9]), tuple(c[10:])
    pt = G[name + "_postype"]
    N = pt.shape[0]
    d_pt = torch.from_numpy(pt).cuda()
    box = gpu.Box.make(L, tilt)
    mesh = gpu.Mesh(*dims, modes)
    mesh.set(1, 1)
    cv = mesh.compute_cv(d_pt, N, box).cpu().item()
    ref_cv, ref_msq = G[name + "_f64_cv"]
    assert mesh.mode_sq() == ref_msq
    rho = G[name + "_f64_rho"]
    assert np.abs(mesh.rho() - rho).max() < 2e-6 * max(1.0, np.abs(rho).max())
    assert cv == pytest.approx(ref_cv, rel=1e-6)
    f = mesh.forces(d_pt, N, box, torch.tensor([bias], dtype=torch.float64, device="cuda")).cpu().numpy()
    fr = G[name + "_f64_force"]
    assert np.abs(f - fr).max() < 2e-4 * np.abs(fr).max()                   # copysignf in the reference's double build
    m = oracle.Mesh(*dims, modes, L, N, "f64", tilt=tilt, literal_copysignf=False)
    m.current_value(pt)
    fo = m.forces(pt, bias)
    assert np.abs(f - fo).max() < 1e-5 * np.abs(fo).max()


def test_mesh_general_path_equals_tiled_path(gpu, oracle, monkeypatch):
    """A power-of-two mesh through both paths (METAD_MESH_GENERAL=1 forces the general one): same density bit for bit when
    the fixed-point scales agree, CV and forces within the tolerances; q_max / virial epilogues of the general path against
    the reference's vectors."""
    import torch
    N, dims, L, modes = 60000, (64, 32, 64), (30.0, 16.0, 31.0), (1.0, -1.0)
    pos, types = rand_pt(N, L, 2, 21)
    d_pt = to_dev(gpu, pos, types)
    box = gpu.Box.make(L)
    bias = torch.tensor([0.4], dtype=torch.float64, device="cuda")
    tiled = gpu.Mesh(*dims, modes)
    tiled.set(1, 1)
    cv_t = tiled.compute_cv(d_pt, N, box).cpu().item()
    f_t = tiled.forces(d_pt, N, box, bias).cpu().numpy()
    monkeypatch.setenv("METAD_MESH_GENERAL", "1")
    gen = gpu.Mesh(*dims, modes)
    gen.set(1, 1)
    cv_g = gen.compute_cv(d_pt, N, box).cpu().item()
    f_g = gen.forces(d_pt, N, box, bias).cpu().numpy()
    assert gen.stats()["rebuilds"] == 0 and tiled.stats()["rebuilds"] == 1          # it really was the other path
    if gen.stats()["fx_scale"] == tiled.stats()["fx_scale"]:
        assert np.array_equal(gen.rho(), tiled.rho())
    assert cv_g == pytest.approx(cv_t, rel=5e-7)
    assert np.abs(f_g - f_t).max() < 5e-6 * np.abs(f_t).max()
    # epilogues of the general path (same checks as test_mesh_qmax_and_virial_against_reference_vectors)
    G = _ref_gold()
    c = G["m1_cfg"]
    dims, L, bias_r, modes = tuple(int(v) for v in c[:3]), tuple(c[3:6]), float(c[6]), tuple(c[7:])
    pt = G["m1_postype"]
    kmin, kmax, n = G["virial_table"]
    kt = np.linspace(kmin, kmax, int(n))
    dK = -2.0 * (kt - 2.0) * np.exp(-(kt - 2.0) ** 2)
    mesh = gpu.Mesh(*dims, modes)
    mesh.set(13, 1)
    mesh.set_table(dK, kmin, kmax)
    cv = mesh.compute_cv(torch.from_numpy(pt).cuda(), pt.shape[0], gpu.Box.make(L)).cpu().item()
    assert mesh.stats()["rebuilds"] == 0
    assert cv == pytest.approx(G["m1_f64_cv"][0], rel=1e-6)
    x = mesh.extras()
    ref_q = G["m1_f64_qmax"]
    assert x["sq_max"] == pytest.approx(ref_q[3], rel=2e-6)
    assert np.allclose(x["q_max"], ref_q[:3], rtol=1e-6, atol=1e-12) or np.allclose(x["q_max"], -ref_q[:3], rtol=1e-6, atol=1e-12)
    ref_v = G["m1_f64_virial"]
    np.testing.assert_allclose(bias_r * x["virial"], ref_v, rtol=2e-5, atol=2e-6 * np.abs(ref_v).max())


def test_mesh_rejects_unsupported(gpu):
    from metadynamics_plugin_b200._abi import MetadError
    with pytest.raises(MetadError, match="1024"):
        gpu.Mesh(2048, 32, 32, [1.0])
    with pytest.raises(MetadError):
        gpu.Mesh(0, 32, 32, [1.0])
    mesh = gpu.Mesh(32, 32, 32, [1.0])
    import torch
    pt = gpu.make_postype(np.zeros((4, 3), np.float32))
    with pytest.raises(MetadError, match="metad_mesh_cv"):
        mesh.forces(pt, 4, gpu.Box.make(5.0), torch.zeros(1, dtype=torch.float64, device="cuda"))


# ------------------------------------------------------------------------------------------------ bias grid
@pytest.mark.parametrize("dims_cfg", [
    dict(cv_min=[-2.0], cv_max=[2.0], num_points=[400], sigma=[0.05]),
    dict(cv_min=[0.0, 0.0], cv_max=[1.0, 2.0], num_points=[20, 30], sigma=[0.25, 0.1]),
    dict(cv_min=[-2.0, 0.0], cv_max=[2.0, 2.0], num_points=[256, 256], sigma=[0.05, 0.1]),
    dict(cv_min=[0.0, -1.0, 2.0], cv_max=[1.0, 1.0, 3.0], num_points=[12, 9, 7], sigma=[0.2, 0.3, 0.25]),
])
@pytest.mark.parametrize("well_tempered", [False, True])
def test_bias_grid_sequence(gpu, oracle, dims_cfg, well_tempered):
    import torch
    d = len(dims_cfg["num_points"])
    kw = dict(W=0.8, T_shift=7.0, T=1.3, stride=3, well_tempered=well_tempered)
    g = gpu.BiasGrid(**dims_cfg, **kw)
    o = oracle.Grid(**dims_cfg, **kw)
    rng = np.random.default_rng(d)
    lo, hi = np.array(dims_cfg["cv_min"]), np.array(dims_cfg["cv_max"])
    s = lo + (hi - lo) * rng.random(d)
    for t in range(14):
        s = np.clip(s + 0.04 * (hi - lo) * rng.normal(size=d), lo - 0.02 * (hi - lo), hi + 0.02 * (hi - lo))   # occasionally off-grid
        if t == 5:
            s = lo + 0.3 * (hi - lo) / (np.array(dims_cfg["num_points"]) - 1)       # forward-difference branch
        if t == 9:
            s = hi - 0.3 * (hi - lo) / (np.array(dims_cfg["num_points"]) - 1)       # backward-difference branch
        b = g.step(t, torch.tensor(s, dtype=torch.float64, device="cuda")).cpu().numpy()
        bo = o.update(t, s)
        np.testing.assert_allclose(b, bo, rtol=1e-9, atol=1e-12)
    for name in ("grid", "reweighted", "weight", "sigma_grid"):
        np.testing.assert_allclose(g.get(name), o.get(name), rtol=1e-10, atol=1e-300, err_msg=name)
    for name in ("hist", "hist_gauss", "hist_delta"):                               # integer grids: bit-exact
        assert np.array_equal(g.get(name), o.get(name).astype(np.uint32)), name
    gs, os_ = g.scalars(), o.scalars()
    assert gs["num_gaussians"] == os_["num_gaussians"] == 5
    assert gs["bias_potential"] == pytest.approx(os_["bias_potential"], rel=1e-10, abs=1e-14)
    assert gs["reweight"] == pytest.approx(os_["reweight"], rel=1e-10)
    assert gs["out_of_bounds"] == os_["out_of_bounds"]


@pytest.mark.parametrize("name", ["g1", "g2", "g3"])
@pytest.mark.parametrize("wt", [0, 1])
def test_bias_grid_against_reference_vectors(gpu, name, wt):
    """metad_grid_step against outputs of the REFERENCE's own IntegratorMetaDynamics.cc (prepRun + updateBiasPotential per
    step; tests/golden/ref_golden.npz): bias factors after every step, final grids, integer histograms bit-exact."""
    import torch
    G = _ref_gold()
    c = G[name + "_cfg"]
    d = int(c[0])
    cfg = dict(cv_min=list(c[1:1 + d]), cv_max=list(c[1 + d:1 + 2 * d]), num_points=[int(v) for v in c[1 + 2 * d:1 + 3 * d]],
               sigma=list(c[1 + 3 * d:1 + 4 * d]))
    g = gpu.BiasGrid(**cfg, W=0.8, T_shift=7.0, T=1.3, stride=3, well_tempered=bool(wt))
    key = "%s_wt%d_" % (name, wt)
    for t, v in enumerate(G[name + "_vals"]):
        b = g.step(t, torch.tensor(v, dtype=torch.float64, device="cuda")).cpu().numpy()
        np.testing.assert_allclose(b, G[key + "bias"][t], rtol=1e-9, atol=1e-12)
    for k in ("grid", "reweighted", "weight", "sigma_grid"):
        np.testing.assert_allclose(g.get(k), G[key + k], rtol=1e-10, atol=1e-300, err_msg=k)
    for k in ("hist", "hist_gauss", "hist_delta"):
        assert np.array_equal(g.get(k), G[key + k]), k
    sc, ref = g.scalars(), G[key + "scalars"]
    assert sc["num_gaussians"] == int(ref[2])
    assert sc["bias_potential"] == pytest.approx(ref[0], rel=1e-10, abs=1e-14) and sc["reweight"] == pytest.approx(ref[1], rel=1e-10)


@pytest.mark.parametrize("well_tempered", [False, True])
def test_bias_grid_multiple_walkers(gpu, oracle, well_tempered):
    """Multiple walkers (IntegratorMetaDynamics.cc:392-410): three walkers with their own CV trajectories share one bias --
    on deposit steps the four delta arrays are summed over the walkers between the Gaussian deposit and the merge.  Device:
    metad_grid_step_deposit / deltas_export -> sum -> deltas_import / metad_grid_step_merge; oracle: the same two halves of
    the restated updateBiasPotential with the sum in between."""
    import torch
    cfg = dict(cv_min=[0.0, 0.0], cv_max=[1.0, 2.0], num_points=[20, 30], sigma=[0.25, 0.1])
    kw = dict(W=0.8, T_shift=7.0, T=1.3, stride=3, well_tempered=well_tempered)
    nw = 3
    gs = [gpu.BiasGrid(**cfg, **kw) for _ in range(nw)]
    os_ = [oracle.Grid(**cfg, **kw) for _ in range(nw)]
    rng = np.random.default_rng(17)
    lo, hi = np.array(cfg["cv_min"]), np.array(cfg["cv_max"])
    s = lo + (hi - lo) * rng.random((nw, 2))
    for t in range(11):
        s = np.clip(s + 0.05 * (hi - lo) * rng.normal(size=(nw, 2)), lo, hi - 1e-9)
        cvs = [torch.tensor(s[k], dtype=torch.float64, device="cuda") for k in range(nw)]
        # device walkers
        for k in range(nw):
            gs[k].step_deposit(t, cvs[k])
        if gs[0].is_deposit_step(t):
            exported = [tuple(x.clone() for x in g.deltas_export()) for g in gs]
            dd = torch.stack([e[0] for e in exported]).sum(0)
            du = torch.stack([e[1] for e in exported]).sum(0).to(torch.int32)
            for g in gs:
                g.deltas_import(dd, du)
        b = [gs[k].step_merge(t, cvs[k]).cpu().numpy().copy() for k in range(nw)]
        # oracle walkers
        for k in range(nw):
            os_[k].update_deposit(t, s[k])
        if t % 3 == 0:
            tot = sum(o.get_deltas() for o in os_)
            for o in os_:
                o.set_deltas(tot)
        bo = [os_[k].update_merge(t, s[k]) for k in range(nw)]
        for k in range(nw):
            np.testing.assert_allclose(b[k], bo[k], rtol=1e-9, atol=1e-12)
    for k in range(nw):
        for name in ("grid", "reweighted", "weight", "sigma_grid"):
            np.testing.assert_allclose(gs[k].get(name), os_[k].get(name), rtol=1e-10, atol=1e-300, err_msg=name)
        for name in ("hist", "hist_gauss", "hist_delta"):
            assert np.array_equal(gs[k].get(name), os_[k].get(name).astype(np.uint32)), name
    # every walker ends with the same bias potential, and it holds all nw x 4 Gaussians
    np.testing.assert_array_equal(gs[0].get("grid"), gs[1].get("grid"))
    assert gs[0].get("hist_gauss").sum() == nw * 4
    # one walker, the two halves back to back == the fused step
    a, b2 = gpu.BiasGrid(**cfg, **kw), gpu.BiasGrid(**cfg, **kw)
    for t in range(5):
        c = torch.tensor(s[0] * (1 - 0.05 * t), dtype=torch.float64, device="cuda")
        ba = a.step(t, c).cpu().numpy().copy()
        bb = b2.step_walkers(t, c, lambda x: None).cpu().numpy().copy()
        np.testing.assert_array_equal(ba, bb)
    np.testing.assert_array_equal(a.get("grid"), b2.get("grid"))


@pytest.mark.parametrize("tag,can", [("ad_all", (1, 1)), ("ad_one", (1, 0))])
def test_bias_grid_adaptive_gaussians_against_reference_vectors(gpu, tag, can):
    """Adaptive Gaussians against the REFERENCE's own integrator (tests/golden/ref_golden.npz, prescribed CV gradients): the
    device sums of products (metad_force_dot), the inverse sigma matrix installed with metad_grid_set_sigma_inv, then the
    bias factors after every step and the final grid / sigma grid."""
    import torch
    G = _ref_gold()
    cfg = dict(cv_min=[0.0, 0.0], cv_max=[1.0, 2.0], num_points=[20, 30], sigma=[0.25, 0.1])
    g = gpu.BiasGrid(**cfg, W=0.8, T_shift=7.0, T=1.3, stride=2, well_tempered=True)
    grads = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in G["ad_grads"]]
    sigma_g = 0.7
    for t, v in enumerate(G["ad_vals"]):
        if t % 2 == 0:
            sq = np.zeros((2, 2))
            for i in range(2):
                for j in range(2):
                    if can[i] and can[j]:
                        sq[i, j] = gpu.force_dot(grads[i], grads[j], sigma_g ** 2).cpu().item()
                    elif i == j:
                        sq[i, j] = cfg["sigma"][i] ** 2
            sinv = np.linalg.inv(np.sqrt(sq))
            np.testing.assert_allclose(sinv, G[tag + "_sigma_inv"][t], rtol=1e-6)      # gradients are float32 products summed in fp64
            g.set_sigma_inv(sinv)
        b = g.step(t, torch.tensor(v, dtype=torch.float64, device="cuda")).cpu().numpy()
        np.testing.assert_allclose(b, G[tag + "_bias"][t], rtol=1e-5, atol=1e-12)
    np.testing.assert_allclose(g.get("grid"), G[tag + "_grid"], rtol=1e-5, atol=1e-300)
    np.testing.assert_allclose(g.get("sigma_grid"), G[tag + "_sigma_grid"], rtol=1e-6, atol=1e-300)


def test_bias_grid_restart_and_flags(gpu, oracle):
    import torch
    cfg = dict(cv_min=[0.0, 0.0], cv_max=[1.0, 2.0], num_points=[20, 30], sigma=[0.25, 0.1])
    g = gpu.BiasGrid(**cfg, stride=1, well_tempered=True)
    o = oracle.Grid(**cfg, stride=1, well_tempered=True)
    dev = lambda v: torch.tensor(v, dtype=torch.float64, device="cuda")
    for t, s in enumerate(([0.1, 1.0], [0.8, 1.0])):
        g.step(t, dev(s)); o.update(t, s)
    r = gpu.BiasGrid(**cfg, stride=1, well_tempered=True)                            # restart from the arrays
    for name in ("grid", "reweighted", "weight", "sigma_grid", "hist", "hist_gauss"):
        r.put(name, g.get(name))
    r.set_num_gaussians(g.scalars()["num_gaussians"])
    b1 = r.step(2, dev([0.4, 1.4])).cpu().numpy()
    b2 = g.step(2, dev([0.4, 1.4])).cpu().numpy()
    np.testing.assert_array_equal(b1, b2)
    np.testing.assert_array_equal(r.get("grid"), g.get("grid"))
    np.testing.assert_allclose(g.get("grid"), (o.update(2, [0.4, 1.4]), o.get("grid"))[1], rtol=1e-10)
    g.set_flags(False, True, 1)                                                     # add_hills = False: no deposit
    before = g.get("grid")
    g.step(3, dev([0.5, 1.0]))
    np.testing.assert_array_equal(g.get("grid"), before)
    assert g.get("hist_delta").sum() == 1
    g.reset_histogram()
    assert g.get("hist").sum() == 0 and g.get("hist_delta").sum() == 0


# ------------------------------------------------------------------------------------------------ umbrella / WTE
def test_umbrella_kinds(gpu, oracle):
    import torch
    for kind, kw in (("harmonic", dict(cv0=0.025, kappa=1.6e7)), ("harmonic", dict(cv0=0.025, kappa=1.6e7, width_flat=0.02)),
                     ("linear", dict(cv0=0.1, scale=3.0)), ("wall", dict(kappa=1.5, scale=0.1)),
                     ("gaussian", dict(cv0=0.1, kappa=0.5, scale=2.0)), ("no_umbrella", {})):
        for val in (0.03, 0.0249, 0.4, -0.2):
            cv = torch.tensor([val], dtype=torch.float64, device="cuda")
            bin_ = torch.tensor([0.37], dtype=torch.float64, device="cuda")
            en = torch.zeros(1, dtype=torch.float64, device="cuda")
            b = gpu.umbrella_apply(kind, cv, bin_, energy_out=en, **kw).cpu().item()
            assert b == pytest.approx(oracle.umbrella_bias(kind, val, 0.37, **kw), rel=1e-12, abs=1e-300)
            assert en.cpu().item() == pytest.approx(oracle.umbrella_potential(kind, val, **kw), rel=1e-12, abs=1e-300)


@pytest.mark.parametrize("N", [1, 1000, 250007])
def test_wte_reduce_and_scale(gpu, oracle, N):
    import torch
    rng = np.random.default_rng(N)
    nf = rng.normal(size=(N, 4)).astype(np.float32)
    tq = rng.normal(size=(N, 4)).astype(np.float32)
    pitch = (N + 15) // 16 * 16
    vir = rng.normal(size=(6 * pitch,)).astype(np.float32)
    d_nf, d_tq, d_vir = (torch.from_numpy(a).cuda() for a in (nf, tq, vir))
    pe = gpu.wte_reduce(d_nf, 1.25).cpu().item()
    assert pe == pytest.approx(oracle.wte_pe(nf, 1.25), rel=1e-12)
    bias = torch.tensor([0.5], dtype=torch.float64, device="cuda")
    gpu.wte_scale(d_nf, d_tq, d_vir, pitch, bias)
    f, t, v, _ = oracle.wte_scale(nf, tq, vir, pitch, 0.5, np.zeros(6))
    np.testing.assert_allclose(d_nf.cpu().numpy(), f, rtol=1e-6)
    np.testing.assert_allclose(d_tq.cpu().numpy(), t, rtol=1e-6)
    np.testing.assert_allclose(d_vir.cpu().numpy(), v, rtol=1e-6)
