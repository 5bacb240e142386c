"""Static evidence about the compiled sm_100a code (no GPU needed): the instructions the design claims are in the library,
and the hot kernels do not spill.  `cuobjdump -sass` / the ptxas logs of the in-tree build are the source, like
profiles/r02_sass_summary.txt (tools/sass_summary.py)."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "metadynamics_plugin_b200", "libmetad_b200.so")
LOG = os.path.join(ROOT, "metadynamics_plugin_b200", "csrc", "_build", "mesh.ptxas.log")


@pytest.fixture(scope="module")
def sass():
    if not shutil.which("cuobjdump") or not os.path.exists(SO):
        pytest.skip("cuobjdump or the built library is not available")
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    kernels, name = {}, None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            kernels[name] = []
        elif name and re.match(r"\s*/\*[0-9a-f]{4,}\*/", line):
            kernels[name].append(line)
    return kernels


def _body(kernels, prefix):
    hits = [k for k in kernels if k.startswith(prefix)]
    assert hits, "no kernel named " + prefix
    return "\n".join(kernels[hits[0]])


def test_tile_movement_uses_the_tma_engine(sass):
    spread = _body(sass, "void metad::mesh::mesh_spread_kernel<4, 10>")          # cache + tensor-map flush: the default at C4
    assert "UTMAREDG.3D.ADD" in spread and "UBLKRED" in spread                   # interior tiles / wrapped tiles
    assert spread.count("ATOMS.ADD") >= 27 and "ATOMS.CAST" not in spread        # native integer atomics, no CAS loop
    gather = _body(sass, "void metad::mesh::mesh_gather_kernel<4, 192, 3, true, false>")
    assert "UTMALDG.3D" in gather and "UBLKCP" in gather and "LDGSTS" in gather
    assert "FFMA2" in gather and "FFMA2" in spread                               # packed fp32 tap arithmetic


def test_triclinic_variants_leave_the_orthorhombic_kernels_alone(sass):
    """The sheared coordinates are evaluated in fp64 (particle_stencil<true>): those instructions live in separate
    instantiations (kSpTri / TRI), the default kernels carry none of them inside their particle loops."""
    ortho = _body(sass, "void metad::mesh::mesh_spread_kernel<4, 10>")
    tri = _body(sass, "void metad::mesh::mesh_spread_kernel<4, 18>")             # cache + triclinic (row flush)
    # (the orthorhombic kernel converts the mode coefficient once per particle for its fp64 sums, nothing else)
    assert tri.count("F2F.F64.F32") >= ortho.count("F2F.F64.F32") + 3 and ortho.count("F2F.F64.F32") <= 1 and ortho.count("F2F.F32.F64") == 0
    assert tri.count("ATOMS.ADD") >= 27 and "ATOMS.CAST" not in tri
    gt = _body(sass, "void metad::mesh::mesh_gather_kernel<4, 192, 3, true, true>")
    assert "UTMALDG.3D" in gt and "FFMA2" in gt


def test_wide_spread_uses_split_32_bit_tiles_and_64_bit_reductions(sass):
    wide = _body(sass, "void metad::mesh::mesh_spread_kernel<4, 6>")              # cache + wide
    assert wide.count("ATOMS.ADD") >= 54 and "ATOMS.CAST" not in wide            # hi / lo parts, native atomics only
    assert re.search(r"(RED|ATOM)G?\.E\.ADD\.64", wide)                          # the flush into the 64-bit mesh: a native 64-bit global add


def test_fused_plane_kernels_use_distributed_shared_memory(sass):
    fwd = _body(sass, "void metad::fft::fft_xy_fwd_kernel<128, 256, 4, 256>")
    assert "UCGABAR_ARV" in fwd and "UCGABAR_WAIT" in fwd                         # cluster barrier (arrive / wait)
    assert re.search(r"\bST\.", fwd)                                              # generic stores into the peer CTA's shared memory


def test_hot_kernels_do_not_spill():
    if not os.path.exists(LOG):
        pytest.skip("ptxas log of the in-tree build not found")
    txt = open(LOG).read()
    want = ["mesh_spread_kernelILi4ELi10E", "mesh_gather_kernelILi4ELi192ELi3ELb1E", "fft_z_fused_kernelILi256ELi3ELb0E",
            "fft_y_kernelILi256ELi1ELi1E", "fft_x_fwd_kernelILi128ELb0E", "fft_x_inv_kernelILi128E", "fft_xy_fwd_kernelILi64ELi128ELi1ELi512E"]
    blocks = txt.split("Compiling entry function '")[1:]
    for w in want:
        hit = [b for b in blocks if w in b.split("'")[0]]
        assert hit, w
        m = re.search(r"(\d+) bytes spill stores", hit[0])
        assert m and int(m.group(1)) == 0, (w, m.group(0) if m else None)
