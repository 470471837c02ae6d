"""Tensor-core pipeline (option gemm=tc: tcgen05 MMAs, TMEM accumulators, TMA-fed weights,
3xTF32 strict mode) against the same oracle and the same tolerances as the FP32-pipe path."""
import os

import numpy as np
import pytest

from conftest import ROOT
from helpers import alg_path, load_golden
from test_gpu_parity import E_ATOL, E_RTOL, F_ATOL, V_RTOL, _random_system, check_outputs, make_pair

pytestmark = pytest.mark.gpu

L1_CASES = ["Cu_r5", "Cu_r15", "aspirin_r15",     # golden cases with l_max = 1 (2, 1, 2 layers)
            "CuPd_r5", "aspirin_r5"]              # and l_max = 2 (3, 2 layers)


@pytest.mark.parametrize("name", L1_CASES)
def test_tc_golden_parity(name, ensure_built):
    atom, lst, z = load_golden(name)
    pair = make_pair(name, z, atom, gemm="tc", keep_edges="1")
    pair.compute(atom, lst)
    assert np.array_equal(pair.handle.get_edges(), z["edge_index"])
    check_outputs(pair, atom, z)
    err_ours = np.abs(atom.f - z["forces64"]).max()
    err_ref = np.abs(z["f"] - z["forces64"]).max()
    assert err_ours < max(10 * err_ref, 2e-5)


@pytest.mark.parametrize("name", ["Cu_r15", "aspirin_r15", "CuPd_r5"])
def test_tc_multichunk(name, ensure_built):
    atom, lst, z = load_golden(name)
    pair = make_pair(name, z, atom, gemm="tc", chunk_edges="4096")
    pair.compute(atom, lst)
    check_outputs(pair, atom, z)


@pytest.mark.parametrize("name", L1_CASES)
def test_tc_intermediates(name, ensure_built):
    from oracle.analytic_numpy import AnalyticAllegro
    from pair_allegro_b200.export import read_alg
    atom, lst, z = load_golden(name)
    pair = make_pair(name, z, atom, gemm="tc", debug="1", keep_edges="1")
    pair.compute(atom, lst)
    h = pair.handle
    hdr, ten = read_alg(alg_path(name))
    ei = z["edge_index"]
    tm = z["type_mapper"]
    zi, zj = tm[atom.type[ei[0]] - 1], tm[atom.type[ei[1]] - 1]
    rvec = atom.x[ei[1]] - atom.x[ei[0]]
    A = AnalyticAllegro(hdr, ten)
    I = A.run(rvec.astype(np.float32).astype(np.float64), ei[0], zi, zj, atom.nlocal)
    E, U = ei.shape[1], 32
    rep, errs = [], {}

    def cmp(label, ours, ref):
        scale = max(1e-6, np.abs(ref).max())
        errs[label] = np.abs(ours - ref).max() / scale
        rep.append("%-10s rel-err %.3e (max|ref| %.3e)" % (label, errs[label], scale))

    has = np.bincount(ei[0], minlength=atom.nlocal) > 0
    for k in range(A.nl):
        cmp("x%d" % k, h.get_output("x%d" % k).reshape(E, 64), I["x"][k])
        g = h.get_output("gamma%d" % k).reshape(-1, A.nsh, U)[:atom.nlocal]
        cmp("gamma%d" % k, g[has], I["Gamma"][k][has])
        if k >= 1:
            cmp("V%d" % k, h.get_output("V%d" % k).reshape(E, U, -1).transpose(0, 2, 1), I["V"][k])
    cmp("edge_energy", h.get_output("edge_energy"), I["e_edge"])
    for k in range(A.nl - 1, -1, -1):
        g = h.get_output("dgamma%d" % k).reshape(-1, A.nsh, U)[:atom.nlocal]
        cmp("dgamma%d" % k, g[has], I["dGamma"][k][has])
    cmp("edge_grad", h.get_output("edge_grad").reshape(E, 3), I["g"])
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "diag_tc_%s.txt" % name), "w") as f:
        f.write("\n".join(rep) + "\n")
    print("\n".join(rep))
    bad = {k: v for k, v in errs.items() if not v < 2e-4}
    assert not bad, bad


@pytest.mark.parametrize("nl,ntypes,lmax", [(1, 1, 1), (2, 2, 1), (3, 3, 1), (1, 2, 2), (2, 1, 2), (3, 2, 2), (2, 2, 3), (3, 3, 3)])
def test_tc_live_oracle(nl, ntypes, lmax, ensure_built, tmp_path):
    from oracle import allegro_torch as AT
    from lmpshim import harness as H
    from oracle.ref_pair import RefPairAllegro
    from pair_allegro_b200.export import export_alg
    from pair_allegro_b200.pair import PairAllegroB200
    names = ["A", "B", "C"][:ntypes]
    pos, types, cell = _random_system(200, 14.0, ntypes, seed=50 + nl)
    atoms = H.make_single_rank(types, pos, cell, [True] * 3, 5.5)
    lst = H.build_full_list(atoms, 5.5)
    cfg = AT.default_config(type_names=names, r_max=4.5, l_max=lmax, num_layers=nl, avg_num_neighbors=20.0,
                            per_type_energy_scales=[1.0 + 0.1 * t for t in range(ntypes)],
                            per_type_energy_shifts=[0.3 * t for t in range(ntypes)], num_bessels=8 if nl != 3 else 12, seed=77 + nl + 10 * lmax)
    pth = str(tmp_path / "m.nequip.pth")
    AT.save_torchscript(cfg, pth)
    export_alg(pth, str(tmp_path / "m.alg"))
    ref = RefPairAllegro()
    ref.coeff(["*", "*", pth] + names, ntypes)
    ref.compute(atoms, lst)
    f_ref, e_ref = atoms.f.copy(), ref.eatom.copy()
    for mode in ("strict", "tf32"):
        atoms.f[:] = 0
        ours = PairAllegroB200(device=0, debug_mode=False)
        ours.coeff(["*", "*", pth] + names, ntypes)
        ours.handle.set_option("gemm", "tc")
        ours.handle.set_option("precision", mode)
        ours.handle.set_option("chunk_edges", "8192")
        ours.compute(atoms, lst)
        nloc = atoms.nlocal
        df = np.abs(atoms.f - f_ref).max()
        de = np.abs(ours.eatom[:nloc] - e_ref[:nloc]).max()
        print("nl=%d mode=%s max|dF|=%.2e max|dE_i|=%.2e (|F|max %.3f)" % (nl, mode, df, de, np.abs(f_ref).max()))
        if mode == "strict":
            np.testing.assert_allclose(ours.eatom[:nloc], e_ref[:nloc], rtol=E_RTOL, atol=E_ATOL)
            assert df < F_ATOL
            assert np.abs(ours.virial - ref.virial).max() < V_RTOL * max(1.0, np.abs(ref.virial).max())
        else:   # fast mode: single TF32 pass, looser stated tolerance
            assert df < 5e-3 * max(1.0, np.abs(f_ref).max()) and de < 5e-3 * max(1.0, np.abs(e_ref).max())


def test_tc_covers_lmax3(ensure_built):
    """l_max = 3 (config C5's architecture) runs on the tensor cores too: one CTA per SM (130 KB of staging), same tolerances"""
    from test_gpu_parity import check_outputs
    atom, lst, z = load_golden("Cu2AgO4_r5")       # l_max = 3, 3 layers
    for pipeline in ("tiled", "fused"):
        atom.f[:] = 0
        pair = make_pair("Cu2AgO4_r5", z, atom, gemm="tc", pipeline=pipeline)
        pair.compute(atom, lst)
        check_outputs(pair, atom, z)
