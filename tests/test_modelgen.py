"""pair_allegro_b200.modelgen (numpy random-init `.alg` writer used by bench.py) and the oracle's
`.alg` loader: both arms of every comparison evaluate bit-identical weights."""
import os

import numpy as np
import pytest

from pair_allegro_b200 import modelgen
from pair_allegro_b200.export import export_alg, read_alg


@pytest.mark.parametrize("L,nl", [(1, 1), (1, 2), (2, 3), (3, 3)])
def test_alg_roundtrip_through_torchscript(L, nl, tmp_path):
    from oracle import allegro_torch as AT
    cfg = modelgen.default_config(l_max=L, num_layers=nl, type_names=["A", "B", "C"],
                                  per_edge_type_cutoff=[[4, 4.5, 5], [4.5, 5, 4], [5, 4, 4.2]],
                                  per_type_energy_scales=[1, 1.1, 1.2], per_type_energy_shifts=[0, .1, .2], seed=5)
    a, p, b = str(tmp_path / "m.alg"), str(tmp_path / "m.nequip.pth"), str(tmp_path / "m2.alg")
    modelgen.random_alg(cfg, a)
    AT.save_torchscript_from_alg(a, p)
    export_alg(p, b)
    h1, t1 = read_alg(a)
    h2, t2 = read_alg(b)
    assert h1 == h2 and list(t1) == list(t2)
    for k in t1:
        assert t1[k].dtype == t2[k].dtype and np.array_equal(t1[k], t2[k]), k


def test_weight_statistics():
    cfg = modelgen.default_config(l_max=2, num_layers=3, seed=3)
    t = modelgen.random_tensors(cfg)
    for k, v in t.items():
        if k.endswith(("alpha", "scales", "shifts", "cutoff_table", "omega")) or k == "readout.w1":
            continue
        assert abs(v.std() * np.sqrt(v.shape[0]) - 1.0) < 0.1, k     # forward-normalised N(0,1)/sqrt(fan_in)
    assert all(0.5 <= float(t["layer%d.alpha" % k][0]) < 1.5 for k in range(3))
    t2 = modelgen.random_tensors(cfg)
    assert all(np.array_equal(t[k], t2[k]) for k in t)               # seeded


@pytest.mark.gpu
@pytest.mark.parametrize("L,nl,gemm", [(1, 2, "tc"), (2, 2, "tc"), (2, 3, "ffma"), (3, 3, "ffma")])
def test_modelgen_model_parity_gpu(L, nl, gemm, ensure_built, tmp_path):
    """bench.py's model source end to end: modelgen `.alg` -> CUDA path vs the same weights in the oracle"""
    from lmpshim import harness as H
    from oracle import allegro_torch as AT
    from oracle.ref_pair import RefPairAllegro
    from pair_allegro_b200.pair import PairAllegroB200
    from test_gpu_parity import E_ATOL, E_RTOL, F_ATOL, V_RTOL, _random_system
    names = ["A", "B"]
    cfg = modelgen.default_config(l_max=L, num_layers=nl, type_names=names, r_max=4.5, avg_num_neighbors=20.0,
                                  per_type_energy_shifts=[0.0, 0.25], seed=11 + L)
    a, p = str(tmp_path / "m.alg"), str(tmp_path / "m.nequip.pth")
    modelgen.random_alg(cfg, a)
    AT.save_torchscript_from_alg(a, p)
    os.rename(a, str(tmp_path / "w.alg"))                    # the reference arm must not depend on the .alg
    pos, types, cell = _random_system(160, 13.0, 2, seed=40 + L)
    atoms = H.make_single_rank(types, pos, cell, [True] * 3, 5.5)
    lst = H.build_full_list(atoms, 5.5)
    ref = RefPairAllegro()
    ref.coeff(["*", "*", p] + names, 2)
    ref.compute(atoms, lst)
    f_ref, e_ref = atoms.f.copy(), ref.eatom.copy()
    atoms.f[:] = 0
    ours = PairAllegroB200(device=0, debug_mode=False)
    ours.coeff(["*", "*", str(tmp_path / "w.alg")] + names, 2)
    ours.handle.set_option("gemm", gemm)
    ours.compute(atoms, lst)
    n = atoms.nlocal
    np.testing.assert_allclose(ours.eatom[:n], e_ref[:n], rtol=E_RTOL, atol=E_ATOL)
    assert np.abs(atoms.f - f_ref).max() < F_ATOL
    assert np.abs(ours.virial - ref.virial).max() < V_RTOL * max(1.0, np.abs(ref.virial).max())
