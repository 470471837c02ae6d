"""GPU parity tests proper: the CUDA path, called through the C-ABI (pair_allegro_b200.pair /
capi -> liballegro_b200.so), against the oracle.

Tolerances (strict fp32 mode, BASELINE.json north_star):
  edge list / neighbour indexing : bit-exact
  per-atom energies              : 1e-5 relative (+1e-5 absolute floor for |E_i| < 1)
  forces                         : 1e-4 eV/A max-abs
  virial                         : 1e-4 relative to max|W| (+1e-4 absolute)
"""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_CASES, ROOT
from helpers import alg_path, golden_config, load_golden

pytestmark = pytest.mark.gpu

E_RTOL, E_ATOL, F_ATOL, V_RTOL = 1e-5, 1e-5, 1e-4, 1e-4


def make_pair(name, z, atom, **opts):
    from pair_allegro_b200.pair import PairAllegroB200
    pair = PairAllegroB200(device=0, debug_mode=False)
    pair.settings([])
    pair.coeff(["*", "*", alg_path(name)] + str(z["type_names"]).split(), atom.ntypes)
    for k, v in opts.items():
        pair.handle.set_option(k, v)
    pair.init_style()
    return pair


def check_outputs(pair, atom, z):
    nl = atom.nlocal
    np.testing.assert_allclose(pair.eatom[:nl], z["eatom"][:nl], rtol=E_RTOL, atol=E_ATOL)
    assert np.abs(atom.f - z["f"]).max() < F_ATOL
    assert abs(pair.eng_vdwl - float(z["eng_vdwl"])) < E_RTOL * max(1.0, np.abs(z["eatom"][:nl]).sum())
    vs = max(1.0, np.abs(z["virial6"]).max())
    assert np.abs(pair.virial - z["virial6"]).max() < V_RTOL * vs
    assert abs(pair.eng_vdwl - pair.eatom[:nl].sum()) < 1e-9 * max(1.0, abs(pair.eng_vdwl))


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_golden_parity(name, ensure_built):
    atom, lst, z = load_golden(name)
    pair = make_pair(name, z, atom, keep_edges="1")
    pair.compute(atom, lst)
    e = pair.handle.get_edges()
    assert e.dtype == np.int64 and np.array_equal(e, z["edge_index"])          # bit-exact
    check_outputs(pair, atom, z)
    # against the fp64-parameter ground truth our error is of the same order as libtorch-fp32's
    err_ours = np.abs(atom.f - z["forces64"]).max()
    err_ref = np.abs(z["f"] - z["forces64"]).max()
    assert err_ours < max(10 * err_ref, 2e-5)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_golden_parity_multichunk(name, ensure_built):
    """same, forcing many small chunks (exercises centre-aligned chunking and tile carries)"""
    atom, lst, z = load_golden(name)
    if z["edge_index"].shape[1] < 300:
        pytest.skip("too few edges to split")
    # Cu_r15 has 1204 edges per atom: chunk must hold at least one full row
    pair = make_pair(name, z, atom, chunk_edges="4096")
    pair.compute(atom, lst)
    check_outputs(pair, atom, z)


@pytest.mark.parametrize("name", ["Cu_r5", "CuPd_r5", "Cu2AgO4_r5", "aspirin_r5", "Cu_r15"])
def test_intermediates(name, ensure_built):
    """per-stage comparison with the fp64 analytic restatement (oracle/analytic_numpy.py);
    a report is written to gpurun_out/ so a failing stage is localised from one GPU run"""
    from oracle.analytic_numpy import AnalyticAllegro
    from pair_allegro_b200.export import read_alg
    atom, lst, z = load_golden(name)
    pair = make_pair(name, z, atom, debug="1", keep_edges="1")
    pair.compute(atom, lst)
    h = pair.handle
    hdr, ten = read_alg(alg_path(name))
    ei = z["edge_index"]
    tm = z["type_mapper"]
    zi, zj = tm[atom.type[ei[0]] - 1], tm[atom.type[ei[1]] - 1]
    rvec = atom.x[ei[1]] - atom.x[ei[0]]
    A = AnalyticAllegro(hdr, ten)
    I = A.run(rvec.astype(np.float32).astype(np.float64), ei[0], zi, zj, atom.nlocal)
    E = ei.shape[1]
    U = 32
    rep = []

    def cmp(label, ours, ref):
        scale = max(1e-6, np.abs(ref).max())
        err = np.abs(ours - ref).max() / scale
        rep.append("%-10s rel-err %.3e (max|ref| %.3e)" % (label, err, scale))
        return err

    errs = {}
    errs["edge_vec"] = cmp("edge_vec", h.get_output("edge_vec").reshape(E, 3), rvec)
    has = np.bincount(ei[0], minlength=atom.nlocal) > 0
    for k in range(A.nl):
        errs["x%d" % k] = cmp("x%d" % k, h.get_output("x%d" % k).reshape(E, 64), I["x"][k])
        g = h.get_output("gamma%d" % k).reshape(-1, A.nsh, U)[:atom.nlocal]
        errs["gamma%d" % k] = cmp("gamma%d" % k, g[has], I["Gamma"][k][has])
        if k >= 1:
            v = h.get_output("V%d" % k).reshape(E, U, -1).transpose(0, 2, 1)
            errs["V%d" % k] = cmp("V%d" % k, v, I["V"][k])
    errs["edge_energy"] = cmp("edge_energy", h.get_output("edge_energy"), I["e_edge"])
    for k in range(A.nl - 1, -1, -1):
        g = h.get_output("dgamma%d" % k).reshape(-1, A.nsh, U)[:atom.nlocal]
        errs["dgamma%d" % k] = cmp("dgamma%d" % k, g[has], I["dGamma"][k][has])
    errs["edge_grad"] = cmp("edge_grad", h.get_output("edge_grad").reshape(E, 3), I["g"])
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "diag_%s.txt" % name), "w") as f:
        f.write("\n".join(rep) + "\n")
    print("\n".join(rep))
    bad = {k: v for k, v in errs.items() if not v < 2e-4}
    assert not bad, bad


def _random_system(n, box, ntypes, seed):
    from lmpshim import harness as H
    rng = np.random.default_rng(seed)
    # jittered lattice so no two atoms are closer than ~1 A
    m = int(np.ceil(n ** (1 / 3)))
    g = np.stack(np.meshgrid(np.arange(m), np.arange(m), np.arange(m), indexing="ij"), -1).reshape(-1, 3)
    pos = (g[rng.permutation(len(g))[:n]] + 0.5) * (box / m) + rng.normal(0, 0.12 * box / m, (n, 3))
    types = rng.integers(1, ntypes + 1, n).astype(np.int32)
    return pos, types, np.eye(3) * box


@pytest.mark.parametrize("L,nl,ntypes", [(1, 1, 1), (1, 2, 1), (2, 2, 2), (2, 3, 2), (3, 3, 4), (3, 1, 2), (1, 3, 3)])
def test_live_oracle_parity(L, nl, ntypes, ensure_built, tmp_path):
    """fresh random-init weights + random periodic box, libtorch (CPU) oracle run on the box"""
    from oracle import allegro_torch as AT
    from lmpshim import harness as H
    from oracle.ref_pair import RefPairAllegro
    from pair_allegro_b200.export import export_alg
    from pair_allegro_b200.pair import PairAllegroB200
    names = ["A", "B", "C", "D"][:ntypes]
    pos, types, cell = _random_system(150, 13.0, ntypes, seed=10 * L + nl)
    r_max = 4.5
    atoms = H.make_single_rank(types, pos, cell, [True] * 3, r_max + 1.0)
    lst = H.build_full_list(atoms, r_max + 1.0)
    cfg = AT.default_config(type_names=names, r_max=r_max, l_max=L, num_layers=nl, avg_num_neighbors=20.0,
                            per_type_energy_scales=[1.0 + 0.1 * t for t in range(ntypes)],
                            per_type_energy_shifts=[0.3 * t for t in range(ntypes)], seed=1000 + L * 7 + nl)
    pth = str(tmp_path / "m.nequip.pth")
    AT.save_torchscript(cfg, pth)
    export_alg(pth, str(tmp_path / "m.alg"))
    ref = RefPairAllegro()
    ref.coeff(["*", "*", pth] + names, ntypes)
    ref.compute(atoms, lst)
    f_ref, e_ref = atoms.f.copy(), ref.eatom.copy()
    atoms.f[:] = 0
    ours = PairAllegroB200(device=0, debug_mode=False)
    ours.coeff(["*", "*", pth] + names, ntypes)          # reference syntax: resolves m.alg next to the .pth
    ours.handle.set_option("keep_edges", "1")
    ours.handle.set_option("chunk_edges", "8192")
    ours.compute(atoms, lst)
    assert np.array_equal(ours.handle.get_edges(), ref.last_input["edge_index"].numpy())
    nloc = atoms.nlocal
    np.testing.assert_allclose(ours.eatom[:nloc], e_ref[:nloc], rtol=E_RTOL, atol=E_ATOL)
    assert np.abs(atoms.f - f_ref).max() < F_ATOL
    assert np.abs(ours.virial - ref.virial).max() < V_RTOL * max(1.0, np.abs(ref.virial).max())
    # newton-on semantics: ghost forces folded back conserve momentum
    ftot = H.reverse_comm_single_rank(atoms, atoms.f)
    assert np.abs(ftot.sum(0)).max() < 1e-3


def test_filter_lt_vs_le(ensure_built):
    """option filter=lt reproduces the Kokkos strict '<' (pair_nequip_allegro_kokkos.cpp:189):
    an atom pair placed exactly at the cutoff is kept by 'le' and dropped by 'lt'."""
    name = "Cu_r5"
    atom, lst, z = load_golden(name)
    # two isolated atoms exactly r_max apart
    class A: pass
    a = A(); a.x = np.array([[0.0, 0, 0], [5.0, 0, 0]]); a.type = np.array([1, 1], dtype=np.int32); a.tag = np.array([1, 2])
    a.nlocal, a.nghost, a.ntypes, a.f = 2, 0, 1, np.zeros((2, 3))
    l = A(); l.inum, l.gnum = 2, 0; l.ilist = np.array([0, 1], dtype=np.int32); l.numneigh = np.array([1, 1], dtype=np.int32)
    l.neigh_flat = np.array([1, 0], dtype=np.int32); l.first = np.array([0, 1], dtype=np.int64)
    for flt, n in (("le", 2), ("lt", 0)):
        pair = make_pair(name, z, a, keep_edges="1", filter=flt)
        pair.compute(a, l)
        assert pair.handle.get_edges().shape[1] == n


def test_empty_domain_and_isolated_atoms(ensure_built):
    name = "aspirin_r5"
    atom, lst, z = load_golden(name)
    cfg = golden_config(z)
    class A: pass
    # nlocal == 0 : silent no-op (pair_nequip_allegro.cpp:341)
    a = A(); a.x = np.zeros((0, 3)); a.type = np.zeros(0, dtype=np.int32); a.tag = np.zeros(0, dtype=np.int64)
    a.nlocal, a.nghost, a.ntypes, a.f = 0, 0, 3, np.zeros((0, 3))
    l = A(); l.inum, l.gnum = 0, 0; l.ilist = np.zeros(0, dtype=np.int32); l.numneigh = np.zeros(0, dtype=np.int32)
    l.neigh_flat = np.zeros(0, dtype=np.int32); l.first = np.zeros(0, dtype=np.int64)
    pair = make_pair(name, z, a)
    pair.compute(a, l)
    assert pair.eng_vdwl == 0.0
    # atoms without any neighbour inside the cutoff: E_i = per-type shift, zero force
    a.x = np.array([[0.0, 0, 0], [20.0, 0, 0], [0, 20.0, 0]]); a.type = np.array([1, 2, 3], dtype=np.int32); a.tag = np.arange(1, 4)
    a.nlocal, a.f = 3, np.zeros((3, 3))
    l.inum = 3; l.ilist = np.arange(3, dtype=np.int32); l.numneigh = np.array([2, 2, 2], dtype=np.int32)
    l.neigh_flat = np.array([1, 2, 0, 2, 0, 1], dtype=np.int32); l.first = np.array([0, 2, 4], dtype=np.int64)
    pair.compute(a, l)
    shifts = np.array(cfg["per_type_energy_shifts"])
    np.testing.assert_allclose(pair.eatom, shifts[z["type_mapper"][a.type - 1]], atol=1e-12)
    assert np.abs(a.f).max() == 0.0


def test_bitwise_determinism(ensure_built):
    """segmented sums in fixed order + fixed-point force accumulation => identical bits run to run"""
    name = "CuPd_r5"
    atom, lst, z = load_golden(name)
    pair = make_pair(name, z, atom, chunk_edges="4096")
    outs = []
    for _ in range(3):
        atom.f[:] = 0
        pair.compute(atom, lst)
        outs.append((atom.f.copy(), pair.eatom.copy(), pair.virial.copy(), pair.eng_vdwl))
    for o in outs[1:]:
        assert np.array_equal(o[0], outs[0][0]) and np.array_equal(o[1], outs[0][1])
        assert np.array_equal(o[2], outs[0][2]) and o[3] == outs[0][3]


def test_device_pointer_entry(ensure_built):
    """alg_compute_device (the Kokkos-twin entry): device-resident x/type/2-D neighbour view, f
    accumulated on the device"""
    name = "CuPd_r5"
    atom, lst, z = load_golden(name)
    pair = make_pair(name, z, atom)
    nl, ntot = atom.nlocal, atom.nlocal + atom.nghost
    maxn = int(lst.numneigh.max())
    nb = np.zeros((nl, maxn), dtype=np.int32)
    for i in range(nl):
        nb[i, :lst.numneigh[i]] = lst.firstneigh(i)
    dev = torch.device("cuda:0")
    d_x = torch.from_numpy(atom.x).to(dev)
    d_type = torch.from_numpy(atom.type).to(dev)
    d_ilist = torch.arange(nl, dtype=torch.int32, device=dev)
    d_num = torch.from_numpy(lst.numneigh[:nl].copy()).to(dev)
    for layout in ("right", "left"):
        if layout == "right":
            d_nb = torch.from_numpy(nb).to(dev); si, sj = maxn, 1
        else:
            d_nb = torch.from_numpy(np.ascontiguousarray(nb.T)).to(dev); si, sj = 1, nl
        d_f = torch.ones(ntot, 3, dtype=torch.float64, device=dev)
        d_e = torch.zeros(ntot, dtype=torch.float64, device=dev)
        torch.cuda.synchronize()
        eng, vir = pair.handle.compute_device(nl, atom.nghost, d_x.data_ptr(), d_type.data_ptr(), d_ilist.data_ptr(),
                                              d_num.data_ptr(), d_nb.data_ptr(), si, sj, d_f.data_ptr(), d_e.data_ptr())
        torch.cuda.synchronize()
        assert np.abs((d_f.cpu().numpy() - 1.0) - z["f"]).max() < F_ATOL
        np.testing.assert_allclose(d_e.cpu().numpy()[:nl], z["eatom"][:nl], rtol=E_RTOL, atol=E_ATOL)
        assert abs(eng - float(z["eng_vdwl"])) < 1e-4
        assert np.abs(vir - z["virial6"]).max() < V_RTOL * max(1.0, np.abs(z["virial6"]).max())


def test_compute_before_type_map_is_an_error(ensure_built):
    from pair_allegro_b200 import capi
    h = capi.Handle(alg_path("Cu_r5"), 0)
    f = np.zeros((1, 3))
    with pytest.raises(capi.AllegroError) as ei:
        h.compute_host(np.zeros((1, 3)), np.ones(1, np.int32), np.zeros(1, np.int32), np.zeros(1, np.int32),
                       np.zeros(0, np.int32), np.zeros(1, np.int64), 1, 0, f)
    assert ei.value.code == -4
