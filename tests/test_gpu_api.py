"""GPU tests of the reference-facing surface: the pybind `_metadynamics` classes and the cv.py / integrate.py API,
driven the way the reference's own scripts (test/test_mesh.py, test/test_2d.py) drive the plugin, checked against
the CPU oracle."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture()
def api():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from metadynamics_plugin_b200 import cv, integrate, hoomd_shim
    hoomd_shim.context.initialize()
    return cv, integrate, hoomd_shim


def test_reference_test_mesh_scenario(api, oracle):
    """reference test/test_mesh.py: N=1000, L=10, cv.mesh(nx=32, mode={'A':1}) under a harmonic umbrella with
    md.integrate.mode_standard; checks cv_mesh, umbrella_energy_mesh and the forces of one step."""
    cv, integrate, hoomd = api
    from metadynamics_plugin_b200 import workloads
    w = workloads.c1()
    pos = w["postype"][:, :3]
    hoomd.init.from_arrays(pos, np.zeros(1000, np.int32), ["A"], 10.0)
    integrate.mode_standard(dt=0.001)
    mesh = cv.mesh(nx=32, mode={'A': 1})
    cv0 = 0.025
    mesh.set_params(umbrella='harmonic', cv0=cv0, kappa=10000 / cv0 ** 2)
    hoomd.run(1)
    o = oracle.Mesh(32, 32, 32, [1.0], 10.0, 1000, "f64", literal_copysignf=False)
    cvo = o.current_value(w["postype"])
    val = mesh.cpp_force.getLogValue("cv_mesh", 1)
    assert val == pytest.approx(cvo, rel=2e-6)                     # float return value of the reference API
    bias = oracle.umbrella_bias("harmonic", float(np.float64(val)), 0.0, cv0=cv0, kappa=10000 / cv0 ** 2)
    f = mesh.get_forces()
    # the device evaluates the umbrella from its fp64 CV; compare force direction/magnitude with the oracle at that bias
    fo = o.forces(w["postype"], bias)
    assert np.abs(f - fo).max() < 2e-4 * np.abs(fo).max()          # kappa (cv - cv0): cancellation amplifies the CV's 1e-6
    e = mesh.cpp_force.getLogValue("umbrella_energy_mesh", 1)
    assert e == pytest.approx(oracle.umbrella_potential("harmonic", float(val), cv0=cv0, kappa=10000 / cv0 ** 2), rel=1e-3)
    assert mesh.cpp_force.getBiasFactor() == 0.0                   # computeForces resets the bias (CollectiveVariable.cc:65)
    assert "cv_mesh" in mesh.cpp_force.getProvidedLogQuantities()
    with pytest.raises(RuntimeError):
        mesh.cpp_force.getLogValue("no_such_quantity", 1)


def test_mesh_cv_in_a_triclinic_box_through_the_api(api, oracle):
    """cv.mesh in a box with tilt factors, through the script-level API: BoxDim carries the tilt to the plan
    (OrderParameterMesh.cc:543, 761-769); the result is the reference's (literal in-cell offsets, see metad_b200.h key 16)."""
    cv, integrate, hoomd = api
    from conftest import triclinic_case
    N, L, tilt = 4000, (11.0, 12.0, 13.0), (0.02, -0.01, 0.03)
    pos, types = triclinic_case(N, L, tilt, 2, 17, faces=False)
    hoomd.init.from_arrays(pos, types, ["A", "B"], L, tilt=tilt)
    integrate.mode_standard(dt=0.001)
    mesh = cv.mesh(nx=32, mode={'A': 1.0, 'B': -1.0})
    cv0 = 0.01
    mesh.set_params(umbrella='harmonic', cv0=cv0, kappa=50.0)
    hoomd.run(1)
    pt = oracle.make_postype(pos, types)
    o = oracle.Mesh(32, 32, 32, [1.0, -1.0], L, N, "f64", tilt=tilt, literal_copysignf=False)
    cvo = o.current_value(pt)
    val = mesh.cpp_force.getLogValue("cv_mesh", 1)
    assert val == pytest.approx(cvo, rel=2e-6)
    bias = oracle.umbrella_bias("harmonic", cvo, 0.0, cv0=cv0, kappa=50.0)
    f = mesh.get_forces()
    fo = o.forces(pt, bias)
    assert np.abs(f - fo).max() < 2e-5 * np.abs(fo).max()          # the umbrella's bias factor carries the CV's 1e-6


def test_mesh_cv_with_a_mesh_that_is_not_a_power_of_two_through_the_api(api, oracle):
    """cv.mesh(nx=24, ny=20, nz=18): the reference takes any mesh size on one rank (OrderParameterMesh.cc:70-79)."""
    cv, integrate, hoomd = api
    N, L = 3000, (11.0, 9.5, 8.0)
    rng = np.random.default_rng(5)
    pos = ((rng.random((N, 3)) - 0.5) * np.asarray(L)).astype(np.float32)
    types = rng.integers(0, 2, N).astype(np.int32)
    hoomd.init.from_arrays(pos, types, ["A", "B"], L)
    integrate.mode_standard(dt=0.001)
    mesh = cv.mesh(nx=24, ny=20, nz=18, mode={'A': 1.0, 'B': -1.0})
    cv0 = 0.01
    mesh.set_params(umbrella='harmonic', cv0=cv0, kappa=50.0)
    hoomd.run(1)
    pt = oracle.make_postype(pos, types)
    o = oracle.Mesh(24, 20, 18, [1.0, -1.0], L, N, "f64", literal_copysignf=False)
    cvo = o.current_value(pt)
    val = mesh.cpp_force.getLogValue("cv_mesh", 1)
    assert val == pytest.approx(cvo, rel=2e-6)
    f = mesh.get_forces()
    fo = o.forces(pt, oracle.umbrella_bias("harmonic", cvo, 0.0, cv0=cv0, kappa=50.0))
    assert np.abs(f - fo).max() < 2e-5 * np.abs(fo).max()


def test_reference_test_2d_scenario(api, oracle, tmp_path):
    """reference test/test_2d.py: one particle, density + aspect-ratio CVs on a 20x30 grid, well-tempered, stride 1,
    grid dumped every step, box rescaled between two run(1) calls; a restart from bias.dat_1 must reproduce
    bias.dat_2 ("identical up to rounding errors")."""
    cv, integrate, hoomd = api
    from metadynamics_plugin_b200 import _metadynamics
    L0 = 10 ** (1. / 3.)
    s = 0.125 ** (1. / 3.)

    def setup():
        hoomd.context.initialize()
        hoomd.init.from_arrays(np.zeros((1, 3), np.float32), [0], ["A"], L0)
        meta = integrate.mode_metadynamics(dt=0.005, mode='well_tempered', stride=1, deltaT=1, W=1)
        density = cv.density(sigma=0.25)
        density.set_grid(cv_min=0, cv_max=1, num_points=20)
        aspect = cv.aspect_ratio(sigma=0.1, dir1=0, dir2=1)
        aspect.set_grid(cv_min=0, cv_max=2, num_points=30)
        return meta

    def rescale():
        pd = hoomd.context.current.system_definition.getParticleData()
        pd.setGlobalBox(_metadynamics.BoxDim(L0 * s, L0 * s, L0 * s))

    meta = setup()
    meta.dump_grid(str(tmp_path / 'bias.dat'), period=1)
    meta.set_params(multiple_walkers=True)
    hoomd.run(1)
    rescale()
    hoomd.run(1)
    grid_final = np.array(meta.cpp_integrator.getGridArray("grid"))
    assert meta.cpp_integrator.getNumGaussians() == 4

    meta2 = setup()
    meta2.restart_from_grid(str(tmp_path / 'bias.dat_1'))
    meta2.dump_grid(str(tmp_path / 'bias_restart.dat'), period=1)
    meta2.set_params(multiple_walkers=True)
    rescale()
    hoomd.run(1)

    a = np.loadtxt(tmp_path / 'bias.dat_2', skiprows=4)
    b = np.loadtxt(tmp_path / 'bias_restart.dat_0', skiprows=4)
    assert a.shape == (600, 8)
    np.testing.assert_allclose(b, a, rtol=2e-9, atol=1e-300)
    hdr = open(tmp_path / 'bias.dat_2').read().splitlines()[:4]
    assert hdr[0] == "#n_cv: 2" and hdr[1] == "#dim:  20 30" and hdr[2] == "#num_gaussians: 4"
    assert hdr[3].split("\t") == ["cv_density", "cv_aspect_ratio", "grid_value", "det_sigma", "num_gaussians", "hist", "hist_reweight", "weight"]

    # the same CV sequence through the oracle: rho = 1/V = 0.1 then 0.8, aspect = 1
    o = oracle.Grid([0.0, 0.0], [1.0, 2.0], [20, 30], [0.25, 0.1], W=1.0, T_shift=1.0, T=1.0, stride=1, well_tempered=True)
    seq =[(0, 1.0 / L0 ** 3), (1, 1.0 / L0 ** 3), (1, 1.0 / (L0 * s) ** 3), (2, 1.0 / (L0 * s) ** 3)]
    for t, rho in seq:
        o.update(t, [float(np.float32(rho)), 1.0])
    np.testing.assert_allclose(grid_final, o.get("grid"), rtol=1e-5, atol=1e-12)      # CV values are float in the API
    assert abs(a[:, 2] - o.get("grid")).max() < 1e-5 * o.get("grid").max()


def test_grid_and_hills_files_match_the_reference_files(api, tmp_path):
    """The same scenario against the files the REFERENCE's own IntegratorMetaDynamics wrote for it
    (tests/golden/test2d_*, generated by tests/golden/make_ref_test2d.py): header lines token for token, integer columns
    exactly, floating columns within the CV's float rounding; and a restart from the reference-written bias.dat_1."""
    cv, integrate, hoomd = api
    from metadynamics_plugin_b200 import _metadynamics
    gold32 = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "test2d_f32")
    gold64 = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "test2d_f64")
    L0 = 10 ** (1. / 3.)
    s = 0.125 ** (1. / 3.)

    def setup(filename=""):
        hoomd.context.initialize()
        hoomd.init.from_arrays(np.zeros((1, 3), np.float32), [0], ["A"], L0)
        meta = integrate.mode_metadynamics(dt=0.005, mode='well_tempered', stride=1, deltaT=1, W=1, filename=filename, overwrite=True)
        cv.density(sigma=0.25).set_grid(cv_min=0, cv_max=1, num_points=20)
        cv.aspect_ratio(sigma=0.1, dir1=0, dir2=1).set_grid(cv_min=0, cv_max=2, num_points=30)
        return meta

    def rescale():
        hoomd.context.current.system_definition.getParticleData().setGlobalBox(_metadynamics.BoxDim(L0 * s, L0 * s, L0 * s))

    def same_file(ours, ref, rtol):
        a, b = open(ours).read().splitlines(), open(ref).read().splitlines()
        assert a[:4] == b[:4] and len(a) == len(b) == 604
        x, y = np.loadtxt(ours, skiprows=4), np.loadtxt(ref, skiprows=4)
        assert np.array_equal(x[:, 4:6], y[:, 4:6])                               # num_gaussians, hist
        for c in (0, 1, 2, 3, 6, 7):
            assert np.abs(x[:, c] - y[:, c]).max() <= rtol * np.abs(y[:, c]).max(), c

    meta = setup(str(tmp_path / "hills.dat"))
    meta.dump_grid(str(tmp_path / 'bias.dat'), period=1)
    hoomd.run(1)
    rescale()
    hoomd.run(1)
    del meta
    for name in ("bias.dat_1", "bias.dat_2"):
        same_file(tmp_path / name, os.path.join(gold32, name), 2e-6)              # reference float build: float grid arithmetic
        same_file(tmp_path / name, os.path.join(gold64, name), 2e-7)              # reference double build: CV values differ by float rounding
    hoomd.context.initialize()                                                    # closes the hills log of the first integrator
    ours, ref = open(tmp_path / "hills.dat").read().splitlines(), open(os.path.join(gold32, "hills.dat")).read().splitlines()
    assert ours[0] == ref[0] and len(ours) == len(ref) == 5
    for a, b in zip(ours[1:], ref[1:]):
        a, b = a.split("\t"), b.split("\t")
        assert a[0] == b[0] and a[3:] == b[3:] and len(a) == len(b) == 6         # timestep; "40", "1", "010": sigma_inv rows without delimiter
        assert float(a[1]) == pytest.approx(float(b[1]), rel=2e-6) and float(a[2]) == pytest.approx(float(b[2]), rel=2e-7)

    meta2 = setup()
    meta2.restart_from_grid(os.path.join(gold64, 'bias.dat_1'))                   # a file written by the reference
    meta2.dump_grid(str(tmp_path / 'bias_restart.dat'), period=1)
    rescale()
    hoomd.run(1)
    same_file(tmp_path / "bias_restart.dat_0", os.path.join(gold64, "bias_restart.dat_0"), 2e-7)


def test_lamellar_metadynamics_steps(api, oracle):
    """cv.lamellar + integrate.mode_metadynamics over a few steps: CV log value, bias factor hand-off, forces, grid."""
    cv, integrate, hoomd = api
    from metadynamics_plugin_b200 import workloads
    w = workloads.c2(N=32768)
    pt = w["postype"]
    types = pt[:, 3].view(np.int32)
    hoomd.init.from_arrays(pt[:, :3], types, ["A", "B"], w["L"])
    meta = integrate.mode_metadynamics(dt=0.005, mode='well_tempered', stride=2, deltaT=7.0, W=1.0)
    lam = cv.lamellar(sigma=0.05, mode=dict(A=1.0, B=-1.0), lattice_vectors=w["lattice_vectors"])
    lam.set_grid(cv_min=-2.0, cv_max=2.0, num_points=400)
    hoomd.run(3)
    cvo, _ = oracle.lamellar_cv(pt, 32768, [1.0, -1.0], w["lattice_vectors"], w["L"])
    assert lam.cpp_force.getLogValue("cv_lamellar", 3) == pytest.approx(cvo, rel=2e-6)
    o = oracle.Grid([-2.0], [2.0], [400], [0.05], W=1.0, T_shift=7.0, T=1.0, stride=2, well_tempered=True)
    for t in (0, 1, 2, 3):                                # prepRun(0) + update(0..2) -> timesteps 0,1,2,3
        bo = o.update(t, [cvo])
    assert meta.cpp_integrator.getNumGaussians() == 2
    np.testing.assert_allclose(np.array(meta.cpp_integrator.getGridArray("grid")), o.get("grid"), rtol=1e-5, atol=1e-12)
    f = lam.get_forces()
    fo = oracle.lamellar_forces(pt, 32768, [1.0, -1.0], w["lattice_vectors"], w["L"], bo[0])
    assert np.abs(f - fo).max() < 1e-3 * np.abs(fo).max() + 1e-12      # dV/ds near a hill centre is ill-conditioned in s
    assert meta.cpp_integrator.getLogValue("bias", 3) == pytest.approx(o.scalars()["bias_potential"], rel=1e-5)
    with pytest.raises(RuntimeError):                     # the set of CVs may not change between runs
        cv.lamellar(sigma=0.05, mode=dict(A=1.0, B=-1.0), lattice_vectors=[(0, 0, 1)], name="x").set_grid(-1, 1, 10)
        hoomd.run(1)


def test_adaptive_gaussians_and_walker_exchange_through_the_api(api, oracle):
    """integrate.mode_metadynamics.set_params(adaptive=True, sigma_g=..., multiple_walkers=True) with two Lamellar CVs: on
    deposit steps the integrator calls computeDerivatives of both CVs, sums the products of their derivatives on the device
    and installs the inverse sigma matrix (IntegratorMetaDynamics.cc:333-341, 1205-1294); the walker hook receives the four
    delta arrays (here: one walker, the hook doubles them -- as if a second identical walker had deposited)."""
    cv, integrate, hoomd = api
    from metadynamics_plugin_b200 import workloads, sharded
    w = workloads.c2(N=16384)
    pt = w["postype"]
    N = pt.shape[0]
    hoomd.init.from_arrays(pt[:, :3], pt[:, 3].view(np.int32), ["A", "B"], w["L"])
    meta = integrate.mode_metadynamics(dt=0.005, mode='well_tempered', stride=2, deltaT=7.0, W=1.0)
    l1 = cv.lamellar(sigma=0.05, mode=dict(A=1.0, B=-1.0), lattice_vectors=[(0, 0, 3)], name="a")
    l2 = cv.lamellar(sigma=0.07, mode=dict(A=1.0, B=-1.0), lattice_vectors=[(0, 3, 0), (3, 0, 0)], name="b")
    l1.set_grid(cv_min=-2.0, cv_max=2.0, num_points=40)
    l2.set_grid(cv_min=-2.0, cv_max=2.0, num_points=30)
    meta.set_params(adaptive=True, sigma_g=0.3, multiple_walkers=True)
    calls = []

    def hook(ptr_d, n_d, ptr_u, n_u):
        import torch
        d, u = sharded.tensor_from_ptr(ptr_d, n_d, torch.float64), sharded.tensor_from_ptr(ptr_u, n_u, torch.int32)
        calls.append((n_d, n_u, float(d[: n_d // 2].sum().item()), int(u.sum().item())))
        d *= 2
        u *= 2
        torch.cuda.synchronize()
    meta.cpp_integrator.walker_allreduce = hook
    hoomd.run(2)                                          # timesteps 0 (prepRun), 1, 2: deposits at 0 and 2
    # derivatives of both CVs from the oracle (force at bias 1), sigma matrix as the reference computes it
    f1 = oracle.lamellar_forces(pt, N, [1.0, -1.0], [(0, 0, 3)], w["L"], 1.0).astype(np.float32)
    f2 = oracle.lamellar_forces(pt, N, [1.0, -1.0], [(0, 3, 0), (3, 0, 0)], w["L"], 1.0).astype(np.float32)
    o = oracle.Grid([-2.0, -2.0], [2.0, 2.0], [40, 30], [0.05, 0.07], W=1.0, T_shift=7.0, T=1.0, stride=2, well_tempered=True)
    sinv = o.compute_sigma([f1, f2], 0.3)
    got = np.array(meta.cpp_integrator.getSigmaInv()).reshape(2, 2)
    if np.all(np.isfinite(sinv)):
        np.testing.assert_allclose(got, sinv, rtol=1e-4)
    else:                                                 # a negative sum of products: NaN in the reference, NaN here
        assert not np.all(np.isfinite(got))
    assert len(calls) == 2 and calls[0][0] == 2 * 40 * 30 and calls[0][1] == 2 * 40 * 30
    assert calls[0][3] == 2                               # hist_delta and hist_gauss_delta of the first deposit: one count each
    assert meta.cpp_integrator.getNumGaussians() == 2
    assert np.array(meta.cpp_integrator.getGridArray("hist_gauss")).sum() == 4       # "two walkers" x two deposits


def test_api_error_behaviour(api):
    cv, integrate, hoomd = api
    hoomd.init.from_arrays(np.zeros((8, 3), np.float32), np.zeros(8, np.int32), ["A", "B"], 5.0)
    with pytest.raises(RuntimeError):
        cv.lamellar(mode=dict(A=1.0), lattice_vectors=[(0, 0, 1)])           # missing mode amplitude for type B
    with pytest.raises(RuntimeError):
        cv.lamellar(mode=dict(A=1.0, B=1.0), lattice_vectors=[])
    with pytest.raises(RuntimeError):
        cv.mesh(mode=[1.0, 1.0], nx=32)                                       # modes must be a dict
    with pytest.raises(RuntimeError, match="power of two"):
        cv.mesh(mode=dict(A=1.0, B=1.0), nx=48)
    with pytest.raises(RuntimeError):
        cv.aspect_ratio(dir1=1, dir2=1)
    with pytest.raises(RuntimeError):
        integrate.mode_metadynamics(dt=0.005, stride=1, mode="flux_tempered")
    m = cv.mesh(mode=dict(A=1.0, B=-1.0), nx=32)
    with pytest.raises(RuntimeError):
        m.set_params(umbrella="parabolic")
    meta = integrate.mode_metadynamics(dt=0.005, stride=1)
    m.set_grid(0.5, 0.1, 10)                                                   # max < min
    with pytest.raises(RuntimeError):
        hoomd.run(1)


def test_potential_energy_cv(api, oracle):
    """cv.potential_energy (well-tempered ensemble): CV = sum of net_force.w; forces = net force scaled by 1+bias."""
    cv, integrate, hoomd = api
    N = 5000
    rng = np.random.default_rng(3)
    sd = hoomd.init.from_arrays(rng.random((N, 3)) - 0.5, np.zeros(N, np.int32), ["A"], 4.0)
    pe = cv.potential_energy(sigma=2.0)
    nf = rng.normal(size=(N, 4)).astype(np.float32)
    pd = sd.getParticleData()
    pd.setNetForce(nf)
    pd.setExternalEnergy(0.5)
    assert pe.cpp_force.requiresNetForce() and not pe.cpp_force.canComputeDerivatives()
    val = pe.cpp_force.getCurrentValue(0)
    assert val == pytest.approx(oracle.wte_pe(nf, 0.5), rel=1e-6)
    pe.cpp_force.setBiasFactor(0.25)
    pe.cpp_force.compute(1)
    np.testing.assert_allclose(pd.getNetForce()[:, :3], nf[:, :3] * 1.25, rtol=1e-6)
    np.testing.assert_array_equal(pd.getNetForce()[:, 3], nf[:, 3])


def test_collective_wrapper(api, oracle):
    """cv.wrap (CollectiveWrapper.cc): CV = potential energy of the wrapped force (sum of force.w + its external energy);
    computeBiasForces scales the wrapped force's own force, torque (all four components) and virial by the bias factor."""
    cv, integrate, hoomd = api
    N = 3000
    rng = np.random.default_rng(12)
    hoomd.init.from_arrays(rng.random((N, 3)) - 0.5, np.zeros(N, np.int32), ["A"], 4.0)
    f4 = rng.normal(size=(N, 4)).astype(np.float32)
    t4 = rng.normal(size=(N, 4)).astype(np.float32)
    vir = rng.normal(size=(6, N)).astype(np.float32)
    pair = hoomd.prescribed_force(f4, t4, vir, external_energy=0.75, name="pair")
    w = cv.wrap(pair, sigma=2.0)
    assert w.cpp_force.getName() == "cv_pair"
    val = w.cpp_force.getCurrentValue(1)
    assert val == pytest.approx(float(f4[:, 3].astype(np.float64).sum()) + 0.75, rel=1e-6)
    assert val == pytest.approx(oracle.wte_pe(f4, 0.75), rel=1e-6)
    w.cpp_force.setBiasFactor(0.4)
    w.cpp_force.compute(1)
    got = pair.cpp_force.getForces()
    np.testing.assert_allclose(got[:, :3], f4[:, :3] * np.float32(0.4), rtol=1e-6)
    np.testing.assert_array_equal(got[:, 3], f4[:, 3])                      # the energy column is not scaled
    np.testing.assert_allclose(pair.cpp_force.getTorques(), t4 * np.float32(0.4), rtol=1e-6)       # torque: x, y, z AND w
    pitch = pair.cpp_force.getVirialPitch()
    np.testing.assert_allclose(pair.cpp_force.getVirial().reshape(6, pitch)[:, :N], vir * np.float32(0.4), rtol=1e-6)
    with pytest.raises(RuntimeError):
        cv.wrap("not a force")


def test_umbrella_reaches_the_host_side_cvs(api, oracle):
    """the factor handed to computeBiasForces is the integrator's dV/ds PLUS the umbrella increment
    (CollectiveVariable.cc:22-66) -- also for the CVs whose bias "force" is a host-side external virial (AspectRatio,
    Density) and for the external-virial part of WellTemperedEnsemble."""
    cv, integrate, hoomd = api
    L = (10.0, 12.0, 9.0)
    sd = hoomd.init.from_arrays(np.zeros((4, 3), np.float32), np.zeros(4, np.int32), ["A"], L)
    integrate.mode_standard(dt=0.001)
    ar = cv.aspect_ratio(dir1=0, dir2=1)
    ar.set_params(umbrella='harmonic', cv0=0.7, kappa=5.0)
    dn = cv.density()
    dn.set_params(umbrella='linear', scale=3.0)
    hoomd.run(1)
    val = oracle.aspect_value(L, 0, 1)
    bias = oracle.umbrella_bias("harmonic", val, 0.0, cv0=0.7, kappa=5.0)
    assert bias != 0.0
    got = np.array([ar.cpp_force.getExternalVirial(i) for i in range(6)])
    np.testing.assert_allclose(got, oracle.aspect_virial(L, 0, 1, bias), rtol=1e-6)
    gotd = np.array([dn.cpp_force.getExternalVirial(i) for i in range(6)])
    np.testing.assert_allclose(gotd, oracle.density_virial(L, 4, oracle.umbrella_bias("linear", oracle.density_value(L, 4), 0.0, scale=3.0)), rtol=1e-6)
    assert np.abs(gotd).max() > 0
    # WellTemperedEnsemble: net force and external virial are scaled by the SAME factor 1 + bias (incl. the umbrella)
    pe = cv.potential_energy(sigma=2.0)
    pe.set_params(umbrella='linear', scale=0.25)
    pd = sd.getParticleData()
    nf = np.arange(16, dtype=np.float32).reshape(4, 4)
    pd.setNetForce(nf)
    for i in range(6):
        pd.setExternalVirial(i, float(i + 1))
    pe.cpp_force.compute(5)
    np.testing.assert_allclose(pd.getNetForce()[:, :3], nf[:, :3] * np.float32(1.25), rtol=1e-6)
    np.testing.assert_allclose([pd.getExternalVirial(i) for i in range(6)], 1.25 * np.arange(1, 7.0), rtol=1e-6)


def test_mesh_log_quantities_and_pressure(api, oracle):
    """cv.mesh log quantities qx_max, qy_max, qz_max, sq_max (OrderParameterMesh.cc:118-122, 1077-1106) and the k-space virial
    that reaches the pressure when the engine asks for it and a kernel table is set (cv.mesh.set_kernel + use_table)."""
    cv, integrate, hoomd = api
    from metadynamics_plugin_b200 import workloads
    N, L = 20000, 20.0
    pos, types = workloads.diblock(N, L, 3, 4)
    sd = hoomd.init.from_arrays(pos, types, ["A", "B"], L)
    integrate.mode_standard(dt=0.001)
    mesh = cv.mesh(nx=32, mode=dict(A=1.0, B=-1.0))
    mesh.set_params(umbrella='linear', scale=0.8)
    assert set(["cv_mesh", "qx_max", "qy_max", "qz_max", "sq_max"]) <= set(mesh.cpp_force.getProvidedLogQuantities())
    pt = oracle.make_postype(pos.astype(np.float32), types)
    o = oracle.Mesh(32, 32, 32, [1.0, -1.0], L, N, "f64")
    o.current_value(pt)
    qo = o.qmax()
    assert mesh.cpp_force.getLogValue("sq_max", 1) == pytest.approx(qo[3], rel=1e-5)
    got = np.array([mesh.cpp_force.getLogValue(k, 1) for k in ("qx_max", "qy_max", "qz_max")])
    assert np.allclose(np.abs(got), np.abs(qo[:3]), rtol=1e-5, atol=1e-9)

    def kernel(k, kmin, kmax, k0):
        return np.exp(-(k - k0) ** 2), -2.0 * (k - k0) * np.exp(-(k - k0) ** 2)
    mesh.set_kernel(kernel, 0.5, 6.0, 64, coeff=dict(k0=2.0))
    mesh.set_params(use_table=True)
    hoomd.run(1)                                   # pressure not asked for: the external virial stays zero (:1062-1071)
    assert all(mesh.cpp_force.getExternalVirial(i) == 0.0 for i in range(6))
    sd.getParticleData().setPressureFlag(True)
    hoomd.run(1)
    kt = np.linspace(0.5, 6.0, 64)
    vo = o.virial(-2.0 * (kt - 2.0) * np.exp(-(kt - 2.0) ** 2), 0.5, 6.0, 0.8)
    got = np.array([mesh.cpp_force.getExternalVirial(i) for i in range(6)])
    assert np.abs(vo).max() > 0
    np.testing.assert_allclose(got, vo, rtol=1e-4, atol=1e-5 * np.abs(vo).max())
