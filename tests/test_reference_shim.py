"""The reference's OWN sources (oracle/_ref, built by oracle/Makefile from
/root/reference/pair_nequip_allegro.cpp against lmpshim + libtorch) pin the oracle
restatement (oracle/ref_pair.py): same TorchScript model file, same LAMMPS-side state ->
same forces / energies / virial / edge dump.  CPU only (CUDA hidden from libtorch)."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from conftest import ROOT
from helpers import golden_config, load_golden

sys.path.insert(0, ROOT)
from lmpshim import driver  # noqa: E402

pytestmark = pytest.mark.skipif(not os.path.exists(driver.REF_LIB), reason="oracle/_ref not built (needs /root/reference)")

SCRIPT = r"""
import os, sys, json
os.environ["CUDA_VISIBLE_DEVICES"] = ""
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import numpy as np
from helpers import load_golden, golden_config
from lmpshim import driver
from oracle import allegro_torch as AT
name = {name!r}
atom, lst, z = load_golden(name)
cfg = golden_config(z)
pth = os.path.join({tmp!r}, name + ".nequip.pth")
AT.save_torchscript(cfg, pth)
lmp = driver.ShimLammps(driver.REF_LIB, atom, lst)
lmp.pair_style([])
lmp.pair_coeff(["*", "*", pth] + str(z["type_names"]).split())
lmp.init(newton_pair=1)
sys.stdout.flush()
out = lmp.compute(eflag=3, vflag=1)
np.savez({tmp!r} + "/out.npz", f=out["f"], eng=out["eng_vdwl"], virial=out["virial"], eatom=out["eatom"],
         cut=lmp.init_one(1, 1), **{{k: v for k, v in lmp.flags().items()}})
"""


def run_reference(name, tmp, debug=False):
    env = dict(os.environ)
    env["CUDA_VISIBLE_DEVICES"] = ""
    if debug:
        env["_NEQUIP_LOG_LEVEL"] = "DEBUG"
    r = subprocess.run([sys.executable, "-c", SCRIPT.format(root=ROOT, name=name, tmp=tmp)], capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    return np.load(os.path.join(tmp, "out.npz")), r.stdout


@pytest.mark.parametrize("name", ["Cu_r5", "Cu2AgO4_r5", "aspirin_r5", "CuPd_r5", "Cu_r15", "aspirin_r15"])
def test_reference_sources_match_oracle_restatement(name):
    atom, lst, z = load_golden(name)
    with tempfile.TemporaryDirectory() as tmp:
        out, _ = run_reference(name, tmp)
    # the real PairNequIPAllegro<false>::compute vs the goldens written by oracle/ref_pair.py
    np.testing.assert_allclose(out["f"], z["f"], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(out["eatom"][:atom.nlocal], z["eatom"][:atom.nlocal], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(float(out["eng"]), float(z["eng_vdwl"]), rtol=1e-6, atol=1e-5)
    np.testing.assert_allclose(out["virial"], z["virial6"], rtol=1e-5, atol=1e-5)
    assert float(out["cut"]) == golden_config(z)["r_max"]                 # init_one (cpp:153-156)
    assert int(out["restartinfo"]) == 0 and int(out["manybody_flag"]) == 1  # cpp:68-69
    assert int(out["neigh_request"]) == 3                                   # REQ_FULL | REQ_GHOST (cpp:146)


@pytest.mark.parametrize("name", ["Cu_r5", "Cu2AgO4_r5", "aspirin_r5", "CuPd_r5", "Cu_r15", "aspirin_r15"])
def test_reference_debug_edge_dump_matches(name):
    """`_NEQUIP_LOG_LEVEL=DEBUG` edge dump of the real reference (cpp:562-565,620-633) == the
    edge list of the oracle (bit-exact indices, printed distances) -- all six fixtures of tests/conftest.py:55-62"""
    atom, lst, z = load_golden(name)
    with tempfile.TemporaryDirectory() as tmp:
        _, stdout = run_reference(name, tmp, debug=True)
    lines = stdout.splitlines()
    a, b = lines.index("Allegro edges: i j rij"), lines.index("end Allegro edges")
    got = [ln.split() for ln in lines[a + 1:b]]
    ei = z["edge_index"]
    assert len(got) == ei.shape[1]
    d = np.linalg.norm(atom.x[ei[0]] - atom.x[ei[1]], axis=1)
    for (i, j, r), e0, e1, rr in zip(got, ei[0], ei[1], d):
        assert int(i) == atom.tag[e0] - 1 and int(j) == atom.tag[e1] - 1
        assert abs(float(r) - rr) < 1e-8                     # printed with 10 significant digits


COMPUTE_SCRIPT = r"""
import os, sys
os.environ["CUDA_VISIBLE_DEVICES"] = ""
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import numpy as np
from helpers import load_golden, golden_config
from lmpshim import driver
from oracle import allegro_torch as AT
name = {name!r}
atom, lst, z = load_golden(name)
pth = os.path.join({tmp!r}, name + ".nequip.pth")
AT.save_torchscript(golden_config(z), pth)
lmp = driver.ShimLammps(driver.REF_LIB, atom, lst)
lmp.set_ghost_owner(atom.owner[atom.nlocal:])
lmp.pair_style([])
lmp.pair_coeff(["*", "*", pth] + str(z["type_names"]).split())
lmp.init(newton_pair=1)
errs = []
for words in (["c", "all", "allegro", "virial"], ["c", "mobile", "allegro", "virial", "9"], ["c", "all", "allegro", "virial", "0"],
              ["c", "all", "allegro/atom", "forces", "3"]):
    try:
        lmp.compute_create(words); errs.append("")
    except driver.ShimError as e:
        errs.append(str(e))
ae = lmp.compute_create(["ae", "all", "allegro/atom", "atomic_energy", "1", "0"])
fo = lmp.compute_create(["fo", "all", "allegro/atom", "forces", "3", "1"])
vi = lmp.compute_create(["vi", "all", "allegro", "virial", "9"])
bad = lmp.compute_create(["bad", "all", "allegro", "virial", "6"])
out = lmp.compute(eflag=3, vflag=1)
try:
    lmp.compute_vector(bad, 6); errs.append("")
except driver.ShimError as e:
    errs.append(str(e))
np.savez({tmp!r} + "/out.npz", f=out["f"], virial=out["virial"], eatom=out["eatom"], errs=np.array(errs),
         c_ae=lmp.compute_peratom(ae, 1), c_fo=lmp.compute_peratom(fo, 3), c_vi=lmp.compute_vector(vi, 9))
"""


def run_reference_compute(name, tmp):
    env = dict(os.environ)
    env["CUDA_VISIBLE_DEVICES"] = ""
    r = subprocess.run([sys.executable, "-c", COMPUTE_SCRIPT.format(root=ROOT, name=name, tmp=tmp)], capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    return np.load(os.path.join(tmp, "out.npz"))


def check_compute_outputs(out, atom):
    """relations every implementation of `compute allegro[/atom]` must satisfy (compute/compute_allegro.cpp)"""
    from lmpshim import harness as H
    n = atom.nlocal
    assert np.array_equal(out["c_ae"][:, 0], out["eatom"][:n])                       # per-atom vector, newton 0: first nlocal rows
    np.testing.assert_allclose(out["c_fo"], H.reverse_comm_single_rank(atom, out["f"]), rtol=0, atol=1e-12)   # newton 1: ghosts folded
    v = out["c_vi"].reshape(3, 3)
    assert np.array_equal(np.array([v[0, 0], v[1, 1], v[2, 2], v[0, 1], v[0, 2], v[1, 2]]), out["virial"])
    e = [str(x) for x in out["errs"]]
    assert "Incorrect args for compute allegro" in e[0]
    assert "can only operate on group 'all'" in e[1]
    assert "Incorrect vector length!" in e[2]
    assert "Incorrect args for compute allegro/atom" in e[3]
    assert "does not match expected 6" in e[4]


def test_reference_compute_allegro():
    """the reference's unmodified compute/compute_allegro.cpp under the same harness"""
    atom, lst, z = load_golden("CuPd_r5")
    with tempfile.TemporaryDirectory() as tmp:
        out = run_reference_compute("CuPd_r5", tmp)
    check_compute_outputs(out, atom)
    np.testing.assert_allclose(out["c_ae"][:, 0], z["eatom"][:atom.nlocal], rtol=1e-5, atol=2e-6)
