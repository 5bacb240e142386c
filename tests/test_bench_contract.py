"""CPU-side checks of bench.py: the reference arm (`--impl reference`) runs without a GPU, prints ONE JSON line with the
contract's keys, describes the same workload as the GPU arm, and times the reference's own classes (oracle/_ref) when they
are built -- falling back to the oracle port, and saying so, when they are not."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_reference_arm(workload, env=None):
    e = dict(os.environ)
    e.update(env or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", workload, "--steps", "2",
                          "--warmup", "1"], capture_output=True, text=True, check=True, env=e, cwd=ROOT).stdout
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1, out
    return json.loads(lines[0])


@pytest.mark.parametrize("workload", ["C1", "C2", "WTE"])
def test_reference_arm_prints_the_contract_line(oracle, workload):
    if workload == "WTE":
        pytest.skip("N = 2^23: minutes on the CPU")
    d = run_reference_arm(workload)
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "cv_bias_force_steps_per_sec" and d["unit"] == "steps/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["gpu_launches"] == 0
    assert d["value"] == pytest.approx(1e3 / d["ms_per_step"])
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] == 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith(workload + ":")
    # same workload description as the GPU arm prints (the driver compares the two config objects)
    sys.path.insert(0, ROOT)
    import bench
    w = bench.make_workload(workload)
    assert d["config"] == bench.config_for(w, 1)


def test_reference_arm_uses_the_references_own_classes_when_built(oracle):
    from oracle import pyref
    if not pyref.available():
        pytest.skip("oracle/_ref is not built (the reference is not mounted and no prebuilt library travelled)")
    d = run_reference_arm("C1")
    assert d["cpu_baseline"]["kind"] == "reference"
    assert set(d["cpu_baseline"]["phases_s_per_step"]) == {"getCurrentValue", "computeBiasForces"}


def test_reference_step_plan_equals_the_oracle(oracle):
    """The persistent object bench.py times (oracle/ref_capi.cc: ref_step_*) computes what the one-shot entry points -- which pin
    the oracle -- compute: CV and forces of the reference's float build."""
    from oracle import pyref
    if not pyref.available():
        pytest.skip("oracle/_ref is not built")
    rng = np.random.default_rng(4)
    N, L = 3000, 9.0
    pos = ((rng.random((N, 3)) - 0.5) * L).astype(np.float32)
    pt = oracle.make_postype(pos, rng.integers(0, 2, N).astype(np.int32))
    plan = pyref.StepPlan("mesh", pt, L, [1.0, -1.0], "f32", dims=(32, 16, 32))
    cv1, _, _ = plan.step(0.7)
    cv2, _, _ = plan.step(0.7)
    one = pyref.mesh((32, 16, 32), [1.0, -1.0], L, pt, 0.7, "f32")
    assert cv1 == cv2 == one["cv"]
    for i in (0, 17, N - 1):
        assert np.array_equal(plan.force(i), one["force"][i])
    lam = pyref.StepPlan("lamellar", pt, L, [1.0, -1.0], "f32", lattice_vectors=[(0, 0, 2), (1, 1, 0)])
    cvl, _, _ = lam.step(-0.3)
    onel = pyref.lamellar([1.0, -1.0], [(0, 0, 2), (1, 1, 0)], L, pt, -0.3, "f32")
    assert cvl == onel["cv"]
    assert np.array_equal(lam.force(5), onel["force"][5])
