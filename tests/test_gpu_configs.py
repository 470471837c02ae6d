"""BASELINE.json configurations C3 (liquid-water-like, 2 species, r_max 6, l_max 2 / 2 layers, virial every
step) and C5 (Li3PO4-like, 3 species, l_max 3 / 3 layers) at sizes the CPU oracle finishes in seconds;
same generators as the full-size boxes (lmpshim/harness.py: water_like_box, multi_species_box)."""
import numpy as np
import pytest

from pair_allegro_b200 import modelgen

pytestmark = pytest.mark.gpu


def _compare(pos, types, cell, names, cfg, gemms, tmp_path, rneigh):
    from lmpshim import harness as H
    from oracle import allegro_torch as AT
    from oracle.ref_pair import RefPairAllegro
    from pair_allegro_b200.pair import PairAllegroB200
    from test_gpu_parity import E_ATOL, E_RTOL, F_ATOL, V_RTOL
    alg, pth = str(tmp_path / "m.alg"), str(tmp_path / "m.nequip.pth")
    modelgen.random_alg(cfg, alg)
    AT.save_torchscript_from_alg(alg, pth)
    atoms = H.make_single_rank(types, pos, cell, [True] * 3, rneigh)
    lst = H.build_full_list(atoms, rneigh)
    ref = RefPairAllegro()
    ref.coeff(["*", "*", pth] + names, len(names))
    ref.compute(atoms, lst)
    f_ref, e_ref, n = atoms.f.copy(), ref.eatom.copy(), atoms.nlocal
    for gemm in gemms:
        atoms.f[:] = 0
        ours = PairAllegroB200(device=0, debug_mode=False)
        ours.coeff(["*", "*", alg] + names, len(names))
        ours.handle.set_option("gemm", gemm)
        ours.handle.set_option("keep_edges", "1")
        ours.handle.set_option("chunk_edges", "16384")
        ours.compute(atoms, lst, eflag=1, vflag=1)
        assert np.array_equal(ours.handle.get_edges(), ref.last_input["edge_index"].numpy())
        np.testing.assert_allclose(ours.eatom[:n], e_ref[:n], rtol=E_RTOL, atol=E_ATOL)
        assert np.abs(atoms.f - f_ref).max() < F_ATOL
        assert np.abs(ours.virial - ref.virial).max() < V_RTOL * max(1.0, np.abs(ref.virial).max())
        assert abs(ours.eng_vdwl - ref.eng_vdwl) < 1e-5 * max(1.0, abs(ref.eng_vdwl))
        print("%s: E %d edges, max|dF| %.2e" % (gemm, ours.handle.get_edges().shape[1], np.abs(atoms.f - f_ref).max()))


def test_c3_water_like(ensure_built, tmp_path):
    from lmpshim import harness as H
    pos, types, cell = H.water_like_box(180, density=0.1, seed=3)           # 540 atoms, box ~17.5 A
    cfg = modelgen.default_config(type_names=["H", "O"], r_max=6.0, l_max=2, num_layers=2, avg_num_neighbors=90.0, seed=3)
    _compare(pos, types, cell, ["H", "O"], cfg, ["tc", "ffma"], tmp_path, 7.0)


def test_c5_multi_species(ensure_built, tmp_path):
    from lmpshim import harness as H
    pos, types, cell = H.multi_species_box(700, fractions=(3, 1, 4), density=0.09, seed=5)   # box ~19.8 A
    cfg = modelgen.default_config(type_names=["Li", "P", "O"], r_max=5.5, l_max=3, num_layers=3, avg_num_neighbors=60.0,
                                  per_edge_type_cutoff=[[5.5, 5.0, 5.5], [5.0, 4.5, 5.0], [5.5, 5.0, 5.5]], seed=5)
    _compare(pos, types, cell, ["Li", "P", "O"], cfg, ["tc", "ffma"], tmp_path, 6.5)
