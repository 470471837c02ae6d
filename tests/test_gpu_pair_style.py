"""The C++ pair style of this repo (src/pair_allegro_b200.cpp, `pair_style allegro`) driven
through the lmpshim harness exactly as the reference's own sources are in
tests/test_reference_shim.py: settings / coeff / init_style / init_one / compute(eflag,vflag),
checked against the oracle goldens (reference test pattern:
/root/reference/tests/test_python_repro_allegro.py:302-355)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN_CASES, ROOT
from helpers import alg_path, golden_config, load_golden

sys.path.insert(0, ROOT)
from lmpshim import driver  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ours_lib(ensure_built):
    if not os.path.exists(driver.OURS_LIB):
        import __graft_entry__ as g
        g.build()
    return driver.OURS_LIB


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_cpp_pair_style_parity(name, ours_lib):
    atom, lst, z = load_golden(name)
    lmp = driver.ShimLammps(ours_lib, atom, lst)
    lmp.pair_style([])
    lmp.pair_coeff(["*", "*", alg_path(name)] + str(z["type_names"]).split())
    lmp.init(newton_pair=1)
    fl = lmp.flags()
    assert fl["restartinfo"] == 0 and fl["manybody_flag"] == 1 and fl["neigh_request"] == 3
    assert lmp.init_one(1, 1) == golden_config(z)["r_max"]
    out = lmp.compute(eflag=3, vflag=1)
    nl = atom.nlocal
    assert np.abs(out["f"] - z["f"]).max() < 1e-4
    np.testing.assert_allclose(out["eatom"][:nl], z["eatom"][:nl], rtol=1e-5, atol=1e-5)
    assert abs(out["eng_vdwl"] - float(z["eng_vdwl"])) < 1e-5 * max(1.0, np.abs(z["eatom"][:nl]).sum())
    assert np.abs(out["virial"] - z["virial6"]).max() < 1e-4 * max(1.0, np.abs(z["virial6"]).max())
    # f is ACCUMULATED (cpp:375-377): a second compute without zeroing doubles it
    out2 = lmp.compute(eflag=1, vflag=0, zero=False)
    np.testing.assert_allclose(out2["f"], 2 * out["f"], rtol=1e-12, atol=1e-12)
    assert np.abs(out2["virial"]).max() == 0.0          # vflag=0: virial untouched after ev_init


def test_cpp_pair_style_errors(ours_lib):
    atom, lst, z = load_golden("Cu_r5")
    lmp = driver.ShimLammps(ours_lib, atom, lst)
    with pytest.raises(driver.ShimError, match="too many arguments"):
        lmp.pair_style(["x"])
    with pytest.raises(driver.ShimError, match="Incorrect args for pair coefficients"):
        lmp.pair_coeff(["*", "*", alg_path("Cu_r5")])
    with pytest.raises(driver.ShimError, match="Only accepts model paths"):
        lmp.pair_coeff(["*", "*", "model.pt", "Cu"])
    with pytest.raises(driver.ShimError, match="cannot open weight file"):
        lmp.pair_coeff(["*", "*", "/nonexistent/m.alg", "Cu"])
    lmp.pair_coeff(["*", "*", alg_path("Cu_r5"), "Cu"])
    with pytest.raises(driver.ShimError, match="requires newton pair on"):
        lmp.init(newton_pair=0)
    lmp.init(newton_pair=1)
    with pytest.raises(driver.ShimError, match="do not support per-atom virial"):
        lmp.compute(eflag=1, vflag=4)


def test_cpp_pair_style_debug_edge_dump(ours_lib):
    """_NEQUIP_LOG_LEVEL=DEBUG prints the reference's edge dump format (cpp:562-565,620-633)"""
    code = r"""
import os, sys
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
from helpers import load_golden, alg_path
from lmpshim import driver
atom, lst, z = load_golden("Cu_r5")
lmp = driver.ShimLammps(driver.OURS_LIB, atom, lst)
lmp.pair_style([]); lmp.pair_coeff(["*", "*", alg_path("Cu_r5"), "Cu"]); lmp.init(1)
sys.stdout.flush()
lmp.compute(3, 1)
""" % (ROOT, ROOT)
    env = dict(os.environ, _NEQUIP_LOG_LEVEL="DEBUG")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    atom, lst, z = load_golden("Cu_r5")
    lines = r.stdout.splitlines()
    a, b = lines.index("Allegro edges: i j rij"), lines.index("end Allegro edges")
    got = [ln.split() for ln in lines[a + 1:b]]
    ei = z["edge_index"]
    assert len(got) == ei.shape[1]
    d = np.linalg.norm(atom.x[ei[0]] - atom.x[ei[1]], axis=1)
    for (i, j, rr), e0, e1, dd in zip(got, ei[0], ei[1], d):
        assert int(i) == atom.tag[e0] - 1 and int(j) == atom.tag[e1] - 1 and abs(float(rr) - dd) < 1e-9


@pytest.mark.parametrize("name", ["CuPd_r5", "aspirin_r15"])
def test_cpp_compute_allegro(name, ours_lib):
    """`compute allegro` / `compute allegro/atom` of this repo (src/compute_allegro_b200.cpp) satisfy the
    same relations -- and print the same errors -- as the reference's unmodified compute
    (tests/test_reference_shim.py::test_reference_compute_allegro), and agree with it numerically"""
    import tempfile

    from test_reference_shim import check_compute_outputs, run_reference_compute
    atom, lst, z = load_golden(name)
    lmp = driver.ShimLammps(ours_lib, atom, lst)
    lmp.set_ghost_owner(atom.owner[atom.nlocal:])
    lmp.pair_style([])
    lmp.pair_coeff(["*", "*", alg_path(name)] + str(z["type_names"]).split())
    lmp.init(newton_pair=1)
    errs = []
    for words in (["c", "all", "allegro", "virial"], ["c", "mobile", "allegro", "virial", "9"], ["c", "all", "allegro", "virial", "0"],
                  ["c", "all", "allegro/atom", "forces", "3"]):
        with pytest.raises(driver.ShimError) as ei:
            lmp.compute_create(words)
        errs.append(str(ei.value))
    ae = lmp.compute_create(["ae", "all", "allegro/atom", "atomic_energy", "1", "0"])
    fo = lmp.compute_create(["fo", "all", "allegro/atom", "forces", "3", "1"])
    vi = lmp.compute_create(["vi", "all", "allegro", "virial", "9"])
    bad = lmp.compute_create(["bad", "all", "allegro", "virial", "6"])
    out = lmp.compute(eflag=3, vflag=1)
    with pytest.raises(driver.ShimError) as ei:
        lmp.compute_vector(bad, 6)
    errs.append(str(ei.value))
    ours = dict(f=out["f"], virial=out["virial"], eatom=out["eatom"], errs=np.array(errs),
                c_ae=lmp.compute_peratom(ae, 1), c_fo=lmp.compute_peratom(fo, 3), c_vi=lmp.compute_vector(vi, 9))
    check_compute_outputs(ours, atom)
    if os.path.exists(driver.REF_LIB):                     # the reference's own compute on the same system (CPU, subprocess)
        with tempfile.TemporaryDirectory() as tmp:
            ref = run_reference_compute(name, tmp)
        np.testing.assert_allclose(ours["c_ae"], ref["c_ae"], rtol=1e-5, atol=1e-5)
        assert np.abs(ours["c_fo"] - ref["c_fo"]).max() < 1e-4
        assert np.abs(ours["c_vi"] - ref["c_vi"]).max() < 1e-4 * max(1.0, np.abs(ref["c_vi"]).max())


def test_neighbour_list_reuse_between_rebuilds(ours_lib):
    """neighbor->ago > 0: the device copy of the list is reused (no flatten + upload), result identical to
    a full upload at the moved positions; ago == 0 or changed atom counts force the upload"""
    from pair_allegro_b200.pair import PairAllegroB200
    name = "CuPd_r5"
    atom, lst, z = load_golden(name)
    rng = np.random.default_rng(1)
    x1 = atom.x + rng.normal(0, 0.02, atom.x.shape)
    x1[atom.nlocal:] = x1[atom.owner[atom.nlocal:]] + (atom.x[atom.nlocal:] - atom.x[atom.owner[atom.nlocal:]])   # ghosts follow owners
    # C++ pair style through the shim
    lmp = driver.ShimLammps(ours_lib, atom, lst)
    lmp.pair_style([])
    lmp.pair_coeff(["*", "*", alg_path(name)] + str(z["type_names"]).split())
    lmp.init(newton_pair=1)
    lmp.compute(eflag=3, vflag=1, neigh_ago=0)
    lmp.set_positions(x1)
    reused = lmp.compute(eflag=3, vflag=1, neigh_ago=1)
    fresh = lmp.compute(eflag=3, vflag=1, neigh_ago=0)
    assert np.array_equal(reused["f"], fresh["f"]) and reused["eng_vdwl"] == fresh["eng_vdwl"]
    assert np.array_equal(reused["eatom"], fresh["eatom"]) and np.array_equal(reused["virial"], fresh["virial"])
    # Python mirror: the stats group tells whether the upload was skipped
    pair = PairAllegroB200(device=0, debug_mode=False)
    pair.coeff(["*", "*", alg_path(name)] + str(z["type_names"]).split(), atom.ntypes)
    pair.compute(atom, lst, neigh_ago=3)                       # nothing cached yet -> uploaded
    assert pair.handle.stats("list_reused", 1)[0] == 0
    pair.compute(atom, lst, neigh_ago=1)
    assert pair.handle.stats("list_reused", 1)[0] == 1
    pair.compute(atom, lst, neigh_ago=0)
    assert pair.handle.stats("list_reused", 1)[0] == 0


@pytest.mark.parametrize("layout_left", [True, False])
@pytest.mark.parametrize("name", ["CuPd_r5", "aspirin_r5", "Cu_r15"])
def test_cpp_pair_style_allegro_kk(name, layout_left, ensure_built):
    """`pair_style allegro/kk` (src/pair_allegro_b200_kokkos.cpp, the twin of the reference's PairAllegroKokkos<false>,
    pair_nequip_allegro_kokkos.cpp:87-353) on device-resident atoms and the KOKKOS 2-D neighbour view in both layouts
    (LayoutLeft = the CUDA default -> thread-per-atom edge build; LayoutRight -> warp-per-atom)"""
    from lmpshim import driver
    atom, lst, z = load_golden(name)
    lmp = driver.ShimLammpsKK(atom, lst, layout_left=layout_left)
    lmp.pair_style([])
    lmp.pair_coeff(["*", "*", alg_path(name)] + str(z["type_names"]).split())
    lmp.init(newton_pair=1)
    for eflag, vflag in ((3, 1), (0, 0), (1, 1)):       # (0, 0): the fully asynchronous call -- forces only
        out = lmp.compute(eflag=eflag, vflag=vflag)
        assert np.abs(out["f"] - z["f"]).max() < 1e-4
        if eflag & 1:
            assert abs(out["eng_vdwl"] - float(z["eng_vdwl"])) < 1e-5 * max(1.0, np.abs(z["eatom"][:atom.nlocal]).sum())
        if eflag & 2:
            np.testing.assert_allclose(out["eatom"][:atom.nlocal], z["eatom"][:atom.nlocal], rtol=1e-5, atol=1e-5)
        if vflag:
            assert np.abs(out["virial"] - z["virial6"]).max() < 1e-4 * max(1.0, np.abs(z["virial6"]).max())
    # the reference rejects `neigh full` for allegro/kk (pair_nequip_allegro_kokkos.cpp:402-405)
    bad = driver.ShimLammpsKK(atom, lst, neighflag=driver.ShimLammpsKK.FULL)
    bad.pair_style([])
    bad.pair_coeff(["*", "*", alg_path(name)] + str(z["type_names"]).split())
    with pytest.raises(driver.ShimError, match="requires the 'neigh half' flag"):
        bad.init(newton_pair=1)
