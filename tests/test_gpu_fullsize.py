"""BASELINE.json's full-size configuration C2 (FCC Ag-like, 63^3 cells = 1,000,188 atoms, ~28.6 M edges)
through size-independent properties, plus an oracle anchor that IS feasible at this size because Allegro
is strictly local: E_i of a sampled centre only needs the atoms inside r_max of it, so the CPU oracle
evaluates one-centre clusters cut out of the big box and must reproduce our per-atom energies."""
import os

import numpy as np
import pytest

from pair_allegro_b200 import modelgen

pytestmark = pytest.mark.gpu

NCELL = int(os.environ.get("ALG_TEST_FULL_NCELL", "63"))
R_MAX, SKIN, LATTICE = 5.0, 1.0, 4.09


@pytest.fixture(scope="module")
def c2(tmp_path_factory):
    from lmpshim import harness as H
    pos, types, cell = H.fcc_box(NCELL, a=LATTICE, jitter=0.05, seed=2)
    atoms = H.make_single_rank(types, pos, cell, [True] * 3, R_MAX + SKIN)
    lst = H.build_full_list(atoms, R_MAX + SKIN)
    d = tmp_path_factory.mktemp("c2")
    alg = str(d / "c2.alg")
    modelgen.random_alg(modelgen.default_config(type_names=["Ag"], r_max=R_MAX, avg_num_neighbors=26.0, seed=2, l_max=1, num_layers=2), alg)
    return atoms, lst, alg


def _run(atoms, lst, alg, **opts):
    from pair_allegro_b200.pair import PairAllegroB200
    pair = PairAllegroB200(device=0, debug_mode=False)
    pair.coeff(["*", "*", alg, "Ag"], 1)
    pair.init_style()
    for k, v in opts.items():
        pair.handle.set_option(k, v)
    atoms.f[:] = 0
    pair.compute(atoms, lst, eflag=1, vflag=1)
    n = atoms.nlocal
    return dict(f=atoms.f.copy(), e=pair.eatom[:n].copy(), pe=pair.eng_vdwl, vir=pair.virial.copy(), edges=int(pair.handle.stats("step", 4)[1]))


def test_c2_full_size_properties(c2, ensure_built, tmp_path):
    from lmpshim import harness as H
    atoms, lst, alg = c2
    n = atoms.nlocal
    if NCELL == 63:
        assert n == 1000188
    tc = _run(atoms, lst, alg, gemm="tc")
    # edge count = independent count of candidate pairs with r^2 <= r_max^2 (f64), the rule of pair_nequip_allegro.cpp:507
    ii = np.repeat(np.arange(n), lst.numneigh[:n])
    d2 = ((atoms.x[ii] - atoms.x[lst.neigh_flat]) ** 2).sum(1)
    assert tc["edges"] == int((d2 <= R_MAX * R_MAX).sum())
    del ii, d2
    # pe == sum of per-atom energies ; Newton's third law after the reverse halo
    assert abs(tc["pe"] - tc["e"].sum()) < 1e-9 * abs(tc["pe"])
    floc = H.reverse_comm_single_rank(atoms, tc["f"])
    assert np.abs(floc.sum(0)).max() < 1e-4
    assert np.abs(tc["f"].sum(0)).max() < 1e-4
    # virial = -sum_edges r (x) dE/dr is symmetric by construction; trace relation with forces on a periodic box:
    # W = sum_i x_i . F_i over locals+ghosts (ghost images carry their shifted x)  == xx+yy+zz
    w = float((atoms.x * tc["f"]).sum())
    assert abs(w - tc["vir"][:3].sum()) < 2e-4 * max(1.0, abs(w))
    # bitwise run-to-run determinism at full size
    tc2 = _run(atoms, lst, alg, gemm="tc")
    assert np.array_equal(tc["f"], tc2["f"]) and np.array_equal(tc["e"], tc2["e"]) and tc["pe"] == tc2["pe"]
    # chunking only regroups per-centre sums: results agree to fp32 round-off
    # (tc = the fused persistent kernel, the default; sm = the chunked edge-tile pipeline in small chunks)
    sm = _run(atoms, lst, alg, gemm="tc", pipeline="tiled", chunk_edges=str(1 << 19))
    assert np.abs(sm["f"] - tc["f"]).max() < 2e-5
    np.testing.assert_allclose(sm["e"], tc["e"], rtol=2e-6, atol=2e-6)
    # the two pipelines (tensor core 3xTF32 / FP32 pipe) agree within the strict tolerances
    ff = _run(atoms, lst, alg, gemm="ffma")
    assert np.abs(ff["f"] - tc["f"]).max() < 1e-4
    np.testing.assert_allclose(ff["e"], tc["e"], rtol=1e-5, atol=1e-5)
    print("total energy: tc %.6f ffma %.6f (rel diff %.2e)" % (tc["pe"], ff["pe"], abs(ff["pe"] - tc["pe"]) / abs(tc["pe"])))
    assert abs(ff["pe"] - tc["pe"]) < 1e-5 * abs(tc["pe"])        # the per-atom bar (1e-5 relative) carried to the sum
    assert np.abs(ff["vir"] - tc["vir"]).max() < 1e-5 * max(1.0, np.abs(tc["vir"]).max())
    # rigid translation (same list): energies and forces unchanged
    x0 = atoms.x.copy()
    atoms.x += np.array([0.37, -1.21, 2.5])
    tr = _run(atoms, lst, alg, gemm="tc")
    atoms.x[:] = x0
    assert np.abs(tr["f"] - tc["f"]).max() < 2e-5
    np.testing.assert_allclose(tr["e"], tc["e"], rtol=5e-6, atol=5e-6)

    # ---- oracle anchor: one-centre clusters cut out of the big box
    from oracle import allegro_torch as AT
    from oracle.ref_pair import RefPairAllegro
    pth = str(tmp_path / "c2.nequip.pth")
    AT.save_torchscript_from_alg(alg, pth)
    ref = RefPairAllegro()
    ref.coeff(["*", "*", pth, "Ag"], 1)
    rng = np.random.default_rng(0)
    worst = 0.0
    for i in rng.choice(n, 48, replace=False):
        nb = lst.firstneigh(int(i))
        x = np.concatenate([atoms.x[i:i + 1], atoms.x[nb]])
        m = len(nb)
        cl = H.Atoms(x=x, type=np.ones(m + 1, dtype=np.int32), tag=np.arange(1, m + 2, dtype=np.int64), nlocal=1, nghost=m, ntypes=1)
        numneigh = np.zeros(m + 1, dtype=np.int32); numneigh[0] = m
        cl_lst = H.NeighList(inum=1, gnum=m, ilist=np.arange(m + 1, dtype=np.int32), numneigh=numneigh,
                             neigh_flat=np.arange(1, m + 1, dtype=np.int32), first=np.zeros(m + 1, dtype=np.int64))
        ref.compute(cl, cl_lst)
        worst = max(worst, abs(ref.eatom[0] - tc["e"][i]) / max(1.0, abs(ref.eatom[0])))
    print("full-size oracle anchor: worst relative E_i error over 48 sampled centres = %.2e" % worst)
    assert worst < 1e-5
