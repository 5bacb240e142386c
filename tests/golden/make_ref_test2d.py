"""Generates tests/golden/test2d_{f64,f32}/ from the REFERENCE's own code: the files (grid dumps, hills log, restart
dumps) that the reference's IntegratorMetaDynamics writes for its test/test_2d.py scenario (oracle/ref_capi.cc
ref_test2d_files: the reference's classes compiled unmodified against the HOOMD stand-in).  They pin the on-disk
formats of writeGrid / readGrid / the hills log (IntegratorMetaDynamics.cc:831-1000, 74-119, 523-550).
Run from the repo root:   python tests/golden/make_ref_test2d.py
"""
import os
import shutil
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyref                     # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
KEEP = {"f64": ["bias.dat_0", "bias.dat_1", "bias.dat_2", "bias_restart.dat_0", "hills.dat"],
        "f32": ["bias.dat_1", "bias.dat_2", "bias_restart.dat_0", "hills.dat"]}

for prec, names in KEEP.items():
    tmp = tempfile.mkdtemp()
    assert pyref.test2d_files(tmp, False, prec) == 4
    assert pyref.test2d_files(tmp, True, prec) == 5
    dst = os.path.join(HERE, "test2d_" + prec)
    os.makedirs(dst, exist_ok=True)
    for n in names:
        shutil.copyfile(os.path.join(tmp, n), os.path.join(dst, n))
    shutil.rmtree(tmp)
    print(prec, names)
