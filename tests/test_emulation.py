"""CPU emulation of the device pipeline: tests/cpu_emul/*.cu run the SAME __host__ __device__ bodies the CUDA
kernels are made of (csrc/mesh_fft.cuh, mesh_fft_kernels.cuh, mesh_kernels.cuh), thread by thread, with the
kernels' tile and index mapping.  This pins the index logic (Stockham stages, R2C packing, the kx=0 plane
untangling, tile-major keys, 27-round spreading, padded-tile merge, gather) without a GPU."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from conftest import build_emul, triclinic_case


def np_chi(dims):
    nx, ny, nz = dims
    nn = lambda n: np.arange(n) < (n // 2 + n % 2)
    return nn(nz)[:, None, None] & nn(ny)[None, :, None] & nn(nx)[None, None, :]


def run_fft(exe, rho, N, mode_sq, stage):
    nz, ny, nx = rho.shape
    with tempfile.TemporaryDirectory() as d:
        fin, fout = os.path.join(d, "in.bin"), os.path.join(d, "out.bin")
        rho.astype(np.float32).tofile(fin)
        subprocess.check_call([exe, str(nx), str(ny), str(nz), repr(float(N)), repr(float(mode_sq)), str(stage), fin, fout])
        raw = np.fromfile(fout, dtype=np.uint8)
    M = rho.size
    return raw[:4 * M].view(np.float32).reshape(nz, ny, nx), raw[4 * M:].view(np.float64)[0]


@pytest.mark.parametrize("dims", [(32, 16, 16), (64, 32, 16), (32, 32, 64), (128, 16, 32), (256, 16, 16), (32, 128, 16),
                                  (32, 16, 256), (512, 16, 16), (1024, 16, 16), (32, 512, 16), (32, 16, 512)])
def test_fft_sweeps_match_numpy(dims):
    exe = build_emul("fft_emul")
    nx, ny, nz = dims
    rng = np.random.default_rng(sum(dims))
    rho = (rng.random((nz, ny, nx)) - 0.5).astype(np.float32)
    # forward x and y sweeps: packed half spectrum
    out, _ = run_fft(exe, rho, 1, 1, 0)
    P = out.view(np.complex64).reshape(nz, ny, nx // 2)
    Fxy = np.fft.fft(np.fft.rfft(rho.astype(np.float64), axis=2), axis=1)
    exp = Fxy[:, :, :nx // 2].copy()
    exp[:, :, 0] = Fxy[:, :, 0] + 1j * Fxy[:, :, nx // 2]          # packed slot: X0 + i X_{nx/2}
    assert np.abs(P - exp).max() < 5e-7 * np.abs(exp).max()
    # full pipeline: CV energy and inverse mesh
    N = rho.size / 4.0
    out, cv = run_fft(exe, rho, N, N, 1)
    f = np.fft.fftn(rho.astype(np.float64)) / N
    a2 = np.abs(f) ** 2
    d = 0.5 * N / N / N
    chi = np_chi(dims)
    s = a2 ** 2 - 2 * d * chi * a2
    s[0, 0, 0] = 0
    assert cv == pytest.approx(0.5 * s.sum(), rel=1e-6)
    inv = (np.fft.ifftn(f * (a2 - d * chi)) * rho.size).real
    assert np.abs(out - inv).max() < 2e-6 * np.abs(inv).max()


def run_mesh(exe, pt, dims, L, n_global, bias, lgT, modes, stale=0.0, tilt=None):
    nx, ny, nz = dims
    N = pt.shape[0]
    env = dict(os.environ)
    env.pop("METAD_EMUL_TILT", None)
    if tilt is not None:
        env["METAD_EMUL_TILT"] = ",".join(repr(float(t)) for t in tilt)
    with tempfile.TemporaryDirectory() as d:
        fin, fout = os.path.join(d, "in.bin"), os.path.join(d, "out.bin")
        pt.tofile(fin)
        subprocess.check_call([exe, str(nx), str(ny), str(nz)] + [repr(float(x)) for x in L] + [str(n_global), repr(bias), str(lgT),
                              repr(float(stale)), str(len(modes))] + [repr(float(m)) for m in modes] + [fin, fout], env=env)
        raw = np.fromfile(fout, dtype=np.uint8)
    M = nx * ny * nz
    cv, msq, shift_err, strays, scale, mismatch = raw[:48].view(np.float64)
    o = 48
    rho = raw[o:o + 4 * M].view(np.float32).reshape(nz, ny, nx); o += 4 * M
    inv = raw[o:o + 4 * M].view(np.float32).reshape(nz, ny, nx); o += 4 * M
    force = raw[o:o + 16 * N].view(np.float32).reshape(N, 4); o += 16 * N
    cells = raw[o:o + 12 * N].view(np.int32).reshape(N, 3)
    return dict(cv=cv, msq=msq, shift_err=shift_err, strays=int(strays), scale=scale, cell_mismatch=int(mismatch), rho=rho, inv=inv,
                force=force, cells=cells)


def mesh_case(N, dims, L, modes, edge, seed):
    rng = np.random.default_rng(seed)
    Lf = np.asarray(L, dtype=np.float64)
    pos = ((rng.random((N, 3)) - 0.5) * Lf).astype(np.float32)
    if edge:      # particles on the upper faces, on the lower corner and one ulp inside
        pos[0] = [np.float32(Lf[0]) / 2, 0, 0]
        pos[1] = [-np.float32(Lf[0]) / 2, np.float32(Lf[1]) / 2, -np.float32(Lf[2]) / 2]
        pos[2] = np.nextafter((Lf / 2).astype(np.float32), np.float32(0))
    return pos, rng.integers(0, len(modes), N)


@pytest.mark.parametrize("N,dims,L,lgT,modes,edge,stale", [
    (1000, (32, 32, 32), (10.0, 10.0, 10.0), 3, (1.0,), False, 0.0),
    (1000, (32, 32, 32), (10.0, 10.0, 10.0), 4, (1.0,), True, 0.9),
    (5000, (32, 16, 64), (10.0, 7.3, 21.1), 3, (1.0, -1.0), True, 0.9),
    (5000, (64, 32, 16), (10.0, 7.3, 21.1), 4, (1.0, -0.5, 2.0), True, 3.5),      # stale order beyond the halo: direct path
    (30000, (32, 32, 32), (31.0, 31.0, 31.0), 3, (1.0,), False, 2.5),
])
def test_mesh_pipeline_matches_oracle(oracle, N, dims, L, lgT, modes, edge, stale):
    exe = build_emul("mesh_emul")
    pos, types = mesh_case(N, dims, L, modes, edge, N + lgT)
    pt = oracle.make_postype(pos, types)
    bias = 0.7
    r = run_mesh(exe, pt, dims, L, N, bias, lgT, modes, stale)
    m = oracle.Mesh(*dims, modes, L, N, "f64", literal_copysignf=False)
    cvo = m.current_value(pt)
    fo = m.forces(pt, bias)
    m32 = oracle.Mesh(*dims, modes, L, N, "f32")
    m32.assign(pt)
    assert np.array_equal(r["cells"], m32.cells())                  # bit-exact against the single-precision build
    assert r["msq"] == m.mode_sq()
    assert r["cell_mismatch"] == 0                                  # conversion-free cell rule == its reference form (C truncation)
    assert r["shift_err"] < 1e-7                                    # compensated fp32 in-cell offset vs its fp64 form
    assert (r["strays"] > 0) == (stale > 2.0)
    assert r["cv"] == pytest.approx(cvo, rel=1e-6)                  # north-star tolerance for CVs
    assert np.abs(r["rho"] - m.mesh).max() < 2e-6 * max(1.0, np.abs(m.mesh).max())
    dinv = r["inv"] - m.inv_re
    dinv -= dinv.mean()                                             # DC removal shifts the inverse mesh by a constant
    assert np.abs(dinv).max() < 5e-6 * np.abs(m.inv_re - m.inv_re.mean()).max()
    assert np.abs(r["force"] - fo).max() < 1e-5 * np.abs(fo).max()  # north-star tolerance for forces
    assert np.all(r["force"][:, 3] == 0)


@pytest.mark.parametrize("literal", [True, False])
@pytest.mark.parametrize("N,dims,L,tilt,lgT,modes,stale", [
    (3000, (32, 32, 32), (10.0, 10.0, 10.0), (0.2, -0.1, 0.15), 3, (1.0,), 0.0),
    (5000, (32, 16, 64), (10.0, 7.3, 21.1), (-0.35, 0.4, 0.25), 3, (1.0, -1.0), 0.4),   # literal offset 12 cells: all weight lost
    (5000, (64, 32, 32), (12.0, 9.0, 10.0), (0.5, 0.0, -0.3), 4, (1.0, -0.5, 2.0), 0.4),
    (4000, (32, 32, 32), (11.0, 12.0, 13.0), (0.0, 0.3, 0.0), 3, (1.0,), 3.5),           # stale order: direct path
    (4000, (32, 32, 32), (11.0, 12.0, 13.0), (0.02, -0.01, 0.03), 3, (1.0, -1.0), 0.4),  # small tilt: offsets below one cell
])
def test_mesh_pipeline_triclinic_matches_oracle(oracle, monkeypatch, N, dims, L, tilt, lgT, modes, stale, literal):
    """Triclinic boxes (BoxDim::makeFraction shears x and y; forces come back through the reciprocal lattice vectors,
    OrderParameterMesh.cc:543-573, 761-769, 836-850).  literal: with the constant the reference's in-cell offsets carry in
    a sheared box (makeFraction(shift + lo), :571-573) -- the behaviour a user of the reference gets; otherwise the
    geometrically correct assignment."""
    exe = build_emul("mesh_emul")
    monkeypatch.setenv("METAD_EMUL_TILT_LITERAL", "1" if literal else "0")
    # With the literal offsets the weights are cut off at |x| = 3/2 and no longer sum to one, so the density is a
    # DISCONTINUOUS function of the cell a particle on a cell face is given to (decided at 1e-16 in the double build):
    # particles within ulps of faces are only meaningful for the continuous (corrected) assignment.
    pos, types = triclinic_case(N, L, tilt, len(modes), N + lgT, faces=not literal)
    pt = oracle.make_postype(pos, types)
    bias = -0.6
    r = run_mesh(exe, pt, dims, L, N, bias, lgT, modes, stale, tilt=tilt)
    m = oracle.Mesh(*dims, modes, L, N, "f64", tilt=tilt, literal_copysignf=False, literal_tilt_offset=literal)
    cvo = m.current_value(pt)
    fo = m.forces(pt, bias)
    m32 = oracle.Mesh(*dims, modes, L, N, "f32", tilt=tilt)
    m32.assign(pt)
    assert np.array_equal(r["cells"], m32.cells())                  # bit-exact against the single-precision build
    # hot stencil vs its fp64 form; the literal offsets (up to 12 cells here) are added in single precision
    assert r["shift_err"] < (1.5e-7 * max(1.0, 8.0 * max(abs(t) for t in tilt) * max(dims)) if literal else 1e-7)
    assert (r["strays"] > 0) == (stale > 2.0)
    # literal offsets of more than a cell drop most of the weight (3 % of it survives in the first case): the fixed-point
    # resolution of a tap is then a larger fraction of the density, and the CV goes with its fourth power
    lost = np.abs(r["rho"].sum()) < 0.5 * N * abs(np.mean(modes)) if literal else False
    assert r["cv"] == pytest.approx(cvo, rel=3e-6 if lost else 1e-6)
    assert np.abs(r["rho"] - m.mesh).max() < 2e-6 * max(1.0, np.abs(m.mesh).max())
    assert np.abs(r["force"] - fo).max() <= 1e-5 * np.abs(fo).max()
    assert np.all(r["force"][:, 3] == 0)


def test_mesh_pipeline_matches_reference_vectors_triclinic():
    """Vector t0 of the reference's own OrderParameterMesh.cc in a triclinic box (double build)."""
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_golden.npz"))
    c = G["t0_cfg"]
    dims, L, tilt, bias, modes = tuple(int(v) for v in c[:3]), tuple(c[3:6]), tuple(c[6:9]), float(c[9]), tuple(c[10:])
    pt = G["t0_postype"]
    r = run_mesh(build_emul("mesh_emul"), pt, dims, L, pt.shape[0], bias, 3, modes, 0.9, tilt=tilt)
    ref_cv, ref_msq = G["t0_f64_cv"]
    assert r["msq"] == ref_msq
    rho = G["t0_f64_rho"]
    assert np.abs(r["rho"] - rho).max() < 2e-6 * max(1.0, np.abs(rho).max())
    assert r["cv"] == pytest.approx(ref_cv, rel=1e-6)
    fr = G["t0_f64_force"]
    assert np.abs(r["force"] - fr).max() < 2e-4 * np.abs(fr).max()


# ---- general path (csrc/mesh_general.cuh): any mesh size
@pytest.mark.parametrize("N,dims,L,modes,edge", [
    (2000, (24, 20, 18), (11.0, 9.5, 8.0), (1.0,), True),              # 4.2.3 | 4.5 | 2.3.3
    (3000, (48, 30, 36), (10.0, 7.3, 21.1), (1.0, -1.0), True),
    (1500, (7, 11, 13), (6.0, 7.0, 8.0), (1.0, -0.5, 2.0), True),      # prime lengths: one radix-n stage
    (1000, (32, 32, 32), (10.0, 10.0, 10.0), (1.0,), True),            # a power of two through the general path
    (500, (3, 1, 2), (4.0, 5.0, 6.0), (1.0,), False),                  # taps alias onto the same cells
    (3000, (100, 6, 45), (30.0, 4.0, 20.0), (1.0, -1.0), False),
    (1000, (1021, 2, 3), (100.0, 3.0, 4.0), (1.0,), False),            # the longest prime line: one stage of 1021 terms per output
    (5000, (16, 16, 16), (8.0, 8.0, 8.0), (1.0,), True),
])
def test_general_mesh_pipeline_matches_oracle(oracle, N, dims, L, modes, edge):
    exe = build_emul("general_emul")
    pos, types = mesh_case(N, dims, L, modes, edge, N + sum(dims))
    pt = oracle.make_postype(pos, types)
    bias = 0.7
    r = run_mesh(exe, pt, dims, L, N, bias, 3, modes, 0.0)
    m = oracle.Mesh(*dims, modes, L, N, "f64", literal_copysignf=False)
    cvo = m.current_value(pt)
    fo = m.forces(pt, bias)
    m32 = oracle.Mesh(*dims, modes, L, N, "f32")
    m32.assign(pt)
    assert np.array_equal(r["cells"], m32.cells())                  # bit-exact against the single-precision build
    assert r["msq"] == m.mode_sq()
    assert r["cv"] == pytest.approx(cvo, rel=1e-6)
    assert np.abs(r["rho"] - m.mesh).max() < 2e-6 * max(1.0, np.abs(m.mesh).max())
    dinv = r["inv"] - m.inv_re
    dinv -= dinv.mean()                                             # k = 0 stays out of the transforms (a constant in the inverse mesh)
    assert np.abs(dinv).max() < 5e-6 * np.abs(m.inv_re - m.inv_re.mean()).max()
    assert np.abs(r["force"] - fo).max() <= 1e-5 * np.abs(fo).max()
    assert np.all(r["force"][:, 3] == 0)


@pytest.mark.parametrize("name", ["t0", "t1", "t2"])
def test_general_mesh_pipeline_matches_reference_vectors(name):
    """Vectors t0-t2 of the reference's own OrderParameterMesh.cc: mesh sizes that are not powers of two (t1: 24 x 20 x 18,
    orthorhombic; t2: 20 x 12 x 18, triclinic) and a triclinic power-of-two mesh (t0), double build."""
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_golden.npz"))
    c = G[name + "_cfg"]
    dims, L, tilt, bias, modes = tuple(int(v) for v in c[:3]), tuple(c[3:6]), tuple(c[6:9]), float(c[9]), tuple(c[10:])
    pt = G[name + "_postype"]
    r = run_mesh(build_emul("general_emul"), pt, dims, L, pt.shape[0], bias, 3, modes, 0.0, tilt=tilt)
    ref_cv, ref_msq = G[name + "_f64_cv"]
    assert r["msq"] == ref_msq
    rho = G[name + "_f64_rho"]
    assert np.abs(r["rho"] - rho).max() < 2e-6 * max(1.0, np.abs(rho).max())
    assert r["cv"] == pytest.approx(ref_cv, rel=1e-6)
    fr = G[name + "_f64_force"]
    assert np.abs(r["force"] - fr).max() < 2e-4 * np.abs(fr).max()              # copysignf in the reference's double build


@pytest.mark.parametrize("literal", [True, False])
def test_general_mesh_pipeline_triclinic(oracle, monkeypatch, literal):
    exe = build_emul("general_emul")
    monkeypatch.setenv("METAD_EMUL_TILT_LITERAL", "1" if literal else "0")
    N, dims, L, tilt, modes = 4000, (36, 20, 30), (11.0, 12.0, 13.0), (0.02, -0.01, 0.03), (1.0, -1.0)
    pos, types = triclinic_case(N, L, tilt, len(modes), 77, faces=not literal)
    pt = oracle.make_postype(pos, types)
    r = run_mesh(exe, pt, dims, L, N, -0.6, 3, modes, 0.0, tilt=tilt)
    m = oracle.Mesh(*dims, modes, L, N, "f64", tilt=tilt, literal_copysignf=False, literal_tilt_offset=literal)
    cvo = m.current_value(pt)
    fo = m.forces(pt, -0.6)
    m32 = oracle.Mesh(*dims, modes, L, N, "f32", tilt=tilt)
    m32.assign(pt)
    assert np.array_equal(r["cells"], m32.cells())
    assert r["cv"] == pytest.approx(cvo, rel=1e-6)
    assert np.abs(r["rho"] - m.mesh).max() < 2e-6 * max(1.0, np.abs(m.mesh).max())
    assert np.abs(r["force"] - fo).max() <= 1e-5 * np.abs(fo).max()


@pytest.mark.parametrize("name", ["m0", "m1", "m2"])
def test_mesh_pipeline_matches_reference_vectors(name):
    """The device code (CPU emulation) against outputs of the REFERENCE's own OrderParameterMesh.cc, double build
    (tests/golden/ref_golden.npz; inputs include particles a few ulps around cell faces).  Forces: the reference's double
    build rounds |x| to float in assignTSCderiv (copysignf), which moves ITS forces by up to 1e-4 of max|F| here."""
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_golden.npz"))
    c = G[name + "_cfg"]
    dims, L, bias, modes = tuple(int(v) for v in c[:3]), tuple(c[3:6]), float(c[6]), tuple(c[7:])
    pt = G[name + "_postype"]
    r = run_mesh(build_emul("mesh_emul"), pt, dims, L, pt.shape[0], bias, 3, modes, 0.9)
    ref_cv, ref_msq = G[name + "_f64_cv"]
    assert r["msq"] == ref_msq and r["cell_mismatch"] == 0
    rho = G[name + "_f64_rho"]
    assert np.abs(r["rho"] - rho).max() < 2e-6 * max(1.0, np.abs(rho).max())
    assert r["cv"] == pytest.approx(ref_cv, rel=1e-6)
    fr = G[name + "_f64_force"]
    assert np.abs(r["force"] - fr).max() < 2e-4 * np.abs(fr).max()


def test_mesh_density_is_order_independent(oracle):
    """Integer accumulation: the density, the CV and the forces are bitwise independent of the tile order."""
    exe = build_emul("mesh_emul")
    dims, L, modes, N = (32, 32, 32), (12.0, 12.0, 12.0), (1.0, -0.7), 4000
    pos, types = mesh_case(N, dims, L, modes, True, 99)
    pt = oracle.make_postype(pos, types)
    ref = run_mesh(exe, pt, dims, L, N, 0.3, 3, modes, 0.0)
    for lgT, stale in ((3, 0.9), (3, 4.0), (4, 0.0), (4, 6.0)):
        r = run_mesh(exe, pt, dims, L, N, 0.3, lgT, modes, stale)
        assert np.array_equal(r["rho"], ref["rho"])
        assert r["cv"] == ref["cv"]
        assert np.array_equal(r["force"], ref["force"])


def test_fixed_point_scale_bounds():
    """fx_scale_for keeps one tap below 2^22 and a full cell below 2^31/4 (host copy of the rule in mesh_kernels.cuh)."""
    for amax, load in ((1.0, 1.0), (1.0, 10.0), (1e-3, 5e-3), (250.0, 4000.0), (1.0, 1e6)):
        tap = 2.0 ** 22 / (0.421875 * amax)
        tot = 2.0 ** 31 / (4 * 5.359375 * max(load, amax))
        lim = min(tap, tot)
        s = 2.0 ** np.floor(np.log2(lim))
        assert s <= lim < 2 * s
        assert 0.421875 * amax * s < 2 ** 22 and 5.359375 * load * s <= 2 ** 29
