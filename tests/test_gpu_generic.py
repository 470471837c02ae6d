"""Width-generic pipeline (csrc/alg_generic.cu, option gemm=generic; the only pipeline for models outside
num_scalar_features=64 / num_tensor_features=32 / MLP 2x64 / readout 32): the golden fixtures at the standard widths, and
models of other widths -- among them the "high-capacity" S=128 / U=64 / H=128 model of BASELINE.json configs[4] -- against
the libtorch (CPU) oracle through the same tolerances as every other parity test."""
import numpy as np
import pytest

from conftest import GOLDEN_CASES
from helpers import load_golden
from pair_allegro_b200 import capi, modelgen
from test_gpu_configs import _compare
from test_gpu_parity import check_outputs, make_pair

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_golden_parity_generic(name, ensure_built):
    atom, lst, z = load_golden(name)
    pair = make_pair(name, z, atom, gemm="generic", keep_edges="1")
    pair.compute(atom, lst)
    assert np.array_equal(pair.handle.get_edges(), z["edge_index"])
    check_outputs(pair, atom, z)
    # many small chunks give the same answer (centre-aligned chunks are independent)
    f1 = atom.f.copy()
    atom.f[:] = 0
    pair2 = make_pair(name, z, atom, gemm="generic", chunk_edges="4096")
    pair2.compute(atom, lst)
    check_outputs(pair2, atom, z)
    assert np.abs(atom.f - f1).max() < 1e-6


WIDTHS = [
    # (l_max, layers, S, U, H, depth, R)
    (1, 2, 128, 64, 128, 2, 64),       # twice the standard widths
    (2, 2, 48, 8, 96, 3, 16),          # nothing a multiple of 64, deeper MLP
    (2, 3, 32, 16, 32, 1, 8),          # narrow, one hidden layer
    (3, 3, 128, 64, 128, 2, 32),       # the high-capacity C5 architecture
]


@pytest.mark.parametrize("L,nl,S,U,H,D,R", WIDTHS)
def test_other_widths_against_oracle(L, nl, S, U, H, D, R, ensure_built, tmp_path):
    from lmpshim import harness as HH
    pos, types, cell = HH.multi_species_box(160 if L == 3 else 260, fractions=(3, 1, 4), density=0.09, seed=7 + L)
    cfg = modelgen.default_config(type_names=["Li", "P", "O"], r_max=4.5, l_max=L, num_layers=nl, avg_num_neighbors=30.0,
                                  num_scalar_features=S, num_tensor_features=U, mlp_width=H, mlp_depth=D, readout_width=R,
                                  per_edge_type_cutoff=[[4.5, 4.0, 4.5], [4.0, 3.5, 4.0], [4.5, 4.0, 4.5]], seed=40 + L + nl)
    _compare(pos, types, cell, ["Li", "P", "O"], cfg, ["generic"], tmp_path, 5.5)


def test_other_widths_select_generic_and_reject_tiled(ensure_built, tmp_path):
    cfg = modelgen.default_config(type_names=["Cu"], l_max=1, num_layers=1, num_scalar_features=128, num_tensor_features=64, mlp_width=128)
    alg = str(tmp_path / "w.alg")
    modelgen.random_alg(cfg, alg)
    h = capi.Handle(alg, 0)
    for v in ("tc", "ffma"):
        with pytest.raises(capi.AllegroError) as ei:
            h.set_option("gemm", v)
        assert "generic" in str(ei.value)
    with pytest.raises(capi.AllegroError):
        h.set_option("pipeline", "fused")
    h.set_option("gemm", "generic")
