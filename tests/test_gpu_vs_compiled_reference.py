"""The CUDA path against the COMPILED reference: the oracle of these tests is literally
`PairNequIPAllegro<false>::compute` (/root/reference/pair_nequip_allegro.cpp:333-407, unmodified, built into
oracle/_ref/libref_pair_allegro.so against the LAMMPS shim + libtorch; it travels to the GPU box prebuilt), run on the
host cores with CUDA hidden, on the same model weights and the same LAMMPS-side state (atoms + full neighbour list).
north_star tolerances: edge list bit-exact (the reference's own DEBUG dump), per-atom energies 1e-5 relative, forces
1e-4 eV/A max-abs, virial 1e-4 of max|W|."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from conftest import GOLDEN_CASES, ROOT
from helpers import alg_path, golden_config, load_golden
from test_gpu_parity import E_ATOL, E_RTOL, F_ATOL, V_RTOL, make_pair

sys.path.insert(0, ROOT)
from lmpshim import driver  # noqa: E402

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.path.exists(driver.REF_LIB), reason="oracle/_ref/libref_pair_allegro.so not built")]

REF_SCRIPT = r"""
import os, sys
os.environ["CUDA_VISIBLE_DEVICES"] = ""
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import numpy as np
from lmpshim import driver, harness as H
from oracle import allegro_torch as AT
z = np.load({inp!r}, allow_pickle=True)
atom = H.Atoms(x=z["x"], type=z["type"], tag=z["tag"], nlocal=int(z["nlocal"]), nghost=int(z["nghost"]), ntypes=int(z["ntypes"]))
lst = H.NeighList(inum=atom.nlocal, gnum=atom.nghost, ilist=z["ilist"], numneigh=z["numneigh"], neigh_flat=z["neigh_flat"], first=z["first"])
pth = os.path.join({tmp!r}, "m.nequip.pth")
AT.save_torchscript_from_alg(str(z["alg"]), pth)          # the SAME weights the CUDA path loads, as the TorchScript file the reference loads
lmp = driver.ShimLammps(driver.REF_LIB, atom, lst)
lmp.pair_style([])
lmp.pair_coeff(["*", "*", pth] + [str(s) for s in z["names"]])
lmp.init(newton_pair=1)
sys.stdout.flush()
out = lmp.compute(eflag=3, vflag=1)
np.savez({tmp!r} + "/ref.npz", f=out["f"], eng=out["eng_vdwl"], virial=out["virial"], eatom=out["eatom"])
"""


def compiled_reference(atom, lst, alg, names, tmp, debug=False):
    inp = os.path.join(tmp, "in.npz")
    ntot = atom.nlocal + atom.nghost
    np.savez(inp, x=atom.x, type=atom.type.astype(np.int32), tag=atom.tag, nlocal=atom.nlocal, nghost=atom.nghost, ntypes=atom.ntypes,
             ilist=lst.ilist[:ntot].astype(np.int32), numneigh=lst.numneigh[:ntot].astype(np.int32), neigh_flat=lst.neigh_flat.astype(np.int32),
             first=lst.first[:ntot].astype(np.int64), alg=alg, names=np.array(names))
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    if debug:
        env["_NEQUIP_LOG_LEVEL"] = "DEBUG"
    r = subprocess.run([sys.executable, "-c", REF_SCRIPT.format(root=ROOT, inp=inp, tmp=tmp)], capture_output=True, text=True, env=env, timeout=1800)
    assert r.returncode == 0, r.stderr[-3000:]
    return np.load(os.path.join(tmp, "ref.npz")), r.stdout


def compare(pair, atom, ref):
    n = atom.nlocal
    np.testing.assert_allclose(pair.eatom[:n], ref["eatom"][:n], rtol=E_RTOL, atol=E_ATOL)
    assert np.abs(atom.f - ref["f"]).max() < F_ATOL
    assert abs(pair.eng_vdwl - float(ref["eng"])) < E_RTOL * max(1.0, np.abs(ref["eatom"][:n]).sum())
    assert np.abs(pair.virial - ref["virial"]).max() < V_RTOL * max(1.0, np.abs(ref["virial"]).max())
    return np.abs(atom.f - ref["f"]).max(), np.abs(pair.eatom[:n] - ref["eatom"][:n]).max()


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_cuda_equals_compiled_reference_on_fixtures(name, ensure_built):
    """all six geometry fixtures of the reference's tests (tests/conftest.py:55-62): outputs AND the DEBUG edge dump"""
    atom, lst, z = load_golden(name)
    names = str(z["type_names"]).split()
    with tempfile.TemporaryDirectory() as tmp:
        ref, stdout = compiled_reference(atom, lst, alg_path(name), names, tmp, debug=True)
    pair = make_pair(name, z, atom, keep_edges="1")
    pair.compute(atom, lst)
    df, de = compare(pair, atom, ref)
    # the reference's own parity hook: "Allegro edges: i j rij" (cpp:562-565, 620-633), tag-1 indices in edge order
    lines = stdout.splitlines()
    a, b = lines.index("Allegro edges: i j rij"), lines.index("end Allegro edges")
    got = np.array([[int(t) for t in ln.split()[:2]] for ln in lines[a + 1:b]], dtype=np.int64).reshape(-1, 2)
    e = pair.handle.get_edges()
    assert got.shape[0] == e.shape[1]
    assert np.array_equal(got[:, 0], atom.tag[e[0]] - 1) and np.array_equal(got[:, 1], atom.tag[e[1]] - 1)     # bit-exact, same order
    print("%s vs compiled reference: max|dF| %.2e eV/A, max|dE_i| %.2e eV, %d edges identical" % (name, df, de, e.shape[1]))


@pytest.mark.parametrize("cfgname", ["c2", "c3", "c5"])
def test_cuda_equals_compiled_reference_on_bench_configs(cfgname, ensure_built, tmp_path):
    """the three architectures bench.py measures (BASELINE.json configs[1], [2], [4]) on oracle-sized boxes of the same
    generators, random-init weights written by the same model generator -- fused pipeline (default) and chunked pipeline"""
    from lmpshim import harness as H
    from pair_allegro_b200 import modelgen
    from pair_allegro_b200.pair import PairAllegroB200
    if cfgname == "c2":
        (pos, types, cell), names, kw, rn = H.fcc_box(6, jitter=0.05, seed=2), ["Ag"], dict(r_max=5.0, l_max=1, num_layers=2, avg_num_neighbors=26.0, seed=2), 6.0
    elif cfgname == "c3":
        (pos, types, cell), names, kw, rn = H.water_like_box(200, seed=3), ["H", "O"], dict(r_max=6.0, l_max=2, num_layers=2, avg_num_neighbors=90.0, seed=3), 7.0
    else:
        (pos, types, cell), names, kw, rn = (H.multi_species_box(600, fractions=(3, 1, 4, 0.5), density=0.09, seed=5), ["Li", "P", "O", "X"],
                                           dict(r_max=5.0, l_max=3, num_layers=3, avg_num_neighbors=47.0, seed=5), 6.0)
    atoms = H.make_single_rank(types, pos, cell, [True] * 3, rn)
    lst = H.build_full_list(atoms, rn)
    alg = str(tmp_path / "m.alg")
    modelgen.random_alg(modelgen.default_config(type_names=names, **kw), alg)
    ref, _ = compiled_reference(atoms, lst, alg, names, str(tmp_path))
    for pipeline in ("fused", "tiled"):
        atoms.f[:] = 0
        pair = PairAllegroB200(device=0, debug_mode=False)
        pair.coeff(["*", "*", alg] + names, len(names))
        pair.init_style()
        pair.handle.set_option("pipeline", pipeline)
        pair.compute(atoms, lst)
        df, de = compare(pair, atoms, ref)
        print("%s %s vs compiled reference: max|dF| %.2e eV/A, max|dE_i| %.2e eV" % (cfgname, pipeline, df, de))
