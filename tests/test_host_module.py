"""CPU-side checks of the host layer (metadynamics_plugin_b200/host -> _metadynamics): the module imports without a GPU,
exports the reference's class surface (reference module.cc:24-41 and the per-class export_* functions), its pure host
pieces reproduce the reference, and anything that needs the device fails loudly instead of falling back."""
import os

import numpy as np
import pytest

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_golden.npz"))


@pytest.fixture(scope="module")
def mod():
    from metadynamics_plugin_b200 import _metadynamics
    return _metadynamics


def test_module_exports_the_reference_class_surface(mod):
    for name in ("CollectiveVariable", "IntegratorMetaDynamics", "LamellarOrderParameter", "LamellarOrderParameterGPU",
                 "OrderParameterMesh", "OrderParameterMeshGPU", "WellTemperedEnsemble", "AspectRatio", "Density", "IndexGrid",
                 "std_vector_int3"):
        assert hasattr(mod, name), name
    # CollectiveVariable.cc:109-130: the umbrella enum and the operator surface
    for name in ("no_umbrella", "linear", "harmonic", "wall", "gaussian"):
        assert hasattr(mod.CollectiveVariable, name), name
    for name in ("getCurrentValue", "setBiasFactor", "computeDerivatives", "canComputeDerivatives", "requiresNetForce",
                 "getUmbrellaPotential", "setUmbrella", "setKappa", "setMinimum", "setWidthFlat", "setScale", "getName"):
        assert hasattr(mod.CollectiveVariable, name), name
    # IntegratorMetaDynamics.cc:1316-1345
    for name in ("registerCollectiveVariable", "removeAllVariables", "isInitialized", "setGrid", "dumpGrid", "restartFromGridFile",
                 "setAddHills", "setMode", "setStride", "setAdaptive", "setSigmaG", "setMultipleWalkers", "resetHistogram",
                 "standard", "well_tempered"):
        assert hasattr(mod.IntegratorMetaDynamics, name), name


def test_indexgrid_host_class_reproduces_the_reference(mod):
    """IndexGrid.cc through the reference's own class (tests/golden/ref_golden.npz, indexgrid_rows)."""
    for row in GOLD["indexgrid_rows"]:
        d = int(row[3])
        lengths, n, idx, coords = [int(v) for v in row[:d]], int(row[4]), int(row[5]), [int(v) for v in row[7:7 + d]]
        g = mod.IndexGrid(lengths)
        assert g.getNumElements() == n
        assert list(g.getCoordinates(idx)) == coords
        assert g.getIndex(coords) == idx


def test_device_objects_fail_loudly_without_a_gpu(mod):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: nothing to refuse")
    with pytest.raises(RuntimeError, match="CUDA"):
        mod.SystemDefinition(10, mod.BoxDim(5, 5, 5), ["A"])
