"""CPU tests of the oracle itself: it must reproduce the committed golden vectors, agree with
the hand-derived analytic chain rule, be O(3)-equivariant, and its edge builder must hit the
known-answer edge counts of the reference's geometry fixtures."""
import os
import tempfile

import numpy as np
import pytest
import torch

from conftest import GOLDEN_CASES
from helpers import alg_path, golden_config, load_golden
from oracle import allegro_torch as AT
from lmpshim import harness as H
from oracle.analytic_numpy import AnalyticAllegro
from oracle.ref_pair import RefPairAllegro
from pair_allegro_b200.export import export_alg, read_alg

EXPECT_EDGES = {"CuPd_r5": 10752, "Cu_r5": 168, "Cu_r15": 4816, "aspirin_r5": 306, "aspirin_r15": 420}


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_reproduces_golden(name):
    """rebuild the TorchScript model from the stored config, run the restated pair style on the
    stored LAMMPS-side state, compare with the stored outputs (bit-exact edges, fp32-level floats)"""
    atom, lst, z = load_golden(name)
    cfg = golden_config(z)
    with tempfile.TemporaryDirectory() as d:
        pth = os.path.join(d, name + ".nequip.pth")
        AT.save_torchscript(cfg, pth)
        pair = RefPairAllegro()
        pair.settings([])
        pair.coeff(["*", "*", pth] + str(z["type_names"]).split(), atom.ntypes)
        pair.init_style()
        pair.compute(atom, lst, loops=lst.numneigh.sum() < 3000)
        # exported weights are identical to the committed ones
        export_alg(pth, os.path.join(d, "m.alg"))
        _, t_new = read_alg(os.path.join(d, "m.alg"))
        _, t_old = read_alg(alg_path(name))
        for k in t_old:
            assert np.array_equal(t_old[k], t_new[k]), k
    assert np.array_equal(pair.last_input["edge_index"].numpy(), z["edge_index"])
    if name in EXPECT_EDGES:
        assert z["edge_index"].shape[1] == EXPECT_EDGES[name]
    # fp32 model: run-to-run summation order in libtorch differs at the 1e-7 level
    np.testing.assert_allclose(atom.f, z["f"], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(pair.eatom, z["eatom"], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(pair.eng_vdwl, float(z["eng_vdwl"]), rtol=1e-6, atol=1e-5)
    np.testing.assert_allclose(pair.virial, z["virial6"], rtol=1e-5, atol=1e-5)
    # pe == sum of per-atom energies (reference test :321)
    assert abs(pair.eng_vdwl - pair.eatom[:atom.nlocal].sum()) < 1e-9


def test_cu2ago4_plain_cutoff_edge_count():
    """the triclinic fixture with a single r_max=5: 262 edges"""
    atom, lst, z = load_golden("Cu2AgO4_r5")
    x = atom.x
    i = np.repeat(np.arange(atom.nlocal), lst.numneigh[:atom.nlocal])
    j = lst.neigh_flat
    d2 = ((x[i] - x[j]) ** 2).sum(1)
    assert int((d2 <= 25.0).sum()) == 262


@pytest.mark.parametrize("L,nl", [(1, 1), (1, 2), (2, 2), (2, 3), (3, 3)])
def test_analytic_backward_matches_autograd(L, nl):
    cfg = AT.default_config(l_max=L, num_layers=nl, avg_num_neighbors=9.0, type_names=["A", "B", "C"],
                            per_type_energy_scales=[1.0, 1.5, 0.7], per_type_energy_shifts=[0.1, -0.2, 0.3],
                            per_edge_type_cutoff=[[5, 4.5, 4], [4.5, 5, 4.2], [4.0, 4.2, 3.8]], seed=7)
    with tempfile.TemporaryDirectory() as d:
        AT.save_torchscript(cfg, d + "/m.nequip.pth")
        export_alg(d + "/m.nequip.pth", d + "/m.alg")
        hdr, ten = read_alg(d + "/m.alg")
    m64 = AT.build_model(cfg, torch.float64)
    g = torch.Generator().manual_seed(3)
    N = 20
    pos = torch.rand(N, 3, dtype=torch.float64, generator=g) * 5
    types = torch.randint(0, 3, (N,), generator=g)
    dd = (pos[None] - pos[:, None]).norm(dim=2)
    cut = m64.cutoff_table[types[:, None], types[None, :]]
    ii, jj = torch.nonzero((dd <= cut) & (dd > 0), as_tuple=True)
    rvec = (pos[jj] - pos[ii]).detach().requires_grad_(True)
    e_edge = m64.edge_energy(rvec, ii, types[ii], types[jj], N)
    e_atom = torch.zeros(N, dtype=torch.float64).index_add(0, ii, e_edge) * m64.inv_sqrt_n
    etot = (e_atom * m64.scales[types] + m64.shifts[types]).sum()
    gt, = torch.autograd.grad(etot, rvec)
    I = AnalyticAllegro(hdr, ten).run(rvec.detach().numpy(), ii.numpy(), types[ii].numpy(), types[jj].numpy(), N)
    assert np.abs(I["e_edge"] - e_edge.detach().numpy()).max() < 1e-12
    assert np.abs(I["g"] - gt.numpy()).max() < 1e-11 * max(1.0, np.abs(gt.numpy()).max())


@pytest.mark.parametrize("L,nl", [(1, 2), (2, 3), (3, 3)])
def test_oracle_equivariance(L, nl):
    """rotation + inversion: energies invariant, forces covariant (pins SH/CG consistency)"""
    cfg = AT.default_config(l_max=L, num_layers=nl, avg_num_neighbors=10.0)
    m = AT.build_model(cfg, torch.float64)
    g = torch.Generator().manual_seed(11)
    N = 16
    pos = torch.rand(N, 3, dtype=torch.float64, generator=g) * 4
    types = torch.randint(0, 2, (N,), generator=g)
    d = (pos[None] - pos[:, None]).norm(dim=2)
    ii, jj = torch.nonzero((d <= 5.0) & (d > 0), as_tuple=True)
    ei = torch.stack([ii, jj])
    q = torch.randn(4, dtype=torch.float64, generator=g)
    q = q / q.norm()
    w, x, y, z = q.tolist()
    R = torch.tensor([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]], dtype=torch.float64)
    o = m(dict(pos=pos, edge_index=ei, atom_types=types))
    for M in (R, -R):
        o2 = m(dict(pos=pos @ M.T, edge_index=ei, atom_types=types))
        assert (o2["atomic_energy"] - o["atomic_energy"]).abs().max() < 1e-12
        assert (o2["forces"] - o["forces"] @ M.T).abs().max() < 1e-12
        assert (o2["virial"][0] - M @ o["virial"][0] @ M.T).abs().max() < 1e-11


def test_forces_are_energy_gradient():
    cfg = AT.default_config(l_max=2, num_layers=2, avg_num_neighbors=10.0)
    m = AT.build_model(cfg, torch.float64)
    g = torch.Generator().manual_seed(5)
    pos = torch.rand(12, 3, dtype=torch.float64, generator=g) * 4
    types = torch.randint(0, 2, (12,), generator=g)
    d = (pos[None] - pos[:, None]).norm(dim=2)
    ii, jj = torch.nonzero((d <= 5.0) & (d > 0), as_tuple=True)
    ei = torch.stack([ii, jj])
    o = m(dict(pos=pos, edge_index=ei, atom_types=types))
    eps = 1e-5
    for k, ax in [(0, 0), (5, 1), (11, 2)]:
        pp, pm = pos.clone(), pos.clone()
        pp[k, ax] += eps
        pm[k, ax] -= eps
        fd = -(m(dict(pos=pp, edge_index=ei, atom_types=types))["atomic_energy"].sum()
               - m(dict(pos=pm, edge_index=ei, atom_types=types))["atomic_energy"].sum()) / (2 * eps)
        assert abs(fd.item() - o["forces"][k, ax].item()) < 1e-7


def test_harness_reverse_comm_conserves_momentum():
    """after ghost forces are folded back onto their owners the net force vanishes (newton on)"""
    atom, lst, z = load_golden("CuPd_r5")
    ftot = H.reverse_comm_single_rank(atom, z["f"])
    assert np.abs(ftot.sum(0)).max() < 1e-5


def test_brick_decomposition_covers_box():
    pos, types, cell = H.fcc_box(6, jitter=0.05, seed=2)
    parts, rank_of, local_index = H.decompose(pos, types, cell, [True] * 3, 4, 6.0)
    assert sum(p.nlocal for p in parts) == len(pos)
    # every ghost is an image of its owner
    L = np.diag(cell)
    for p in parts:
        gh = slice(p.nlocal, p.nlocal + p.nghost)
        own_pos = np.stack([parts[r].x[i] for r, i in zip(p.owner_rank[gh], p.owner_index[gh])])
        d = p.x[gh] - own_pos
        assert np.abs(d - np.round(d / L) * L).max() < 1e-9


@pytest.mark.parametrize("seed,box,rcut", [(0, (7.0, 8.0, 9.0), 5.0), (1, (4.0, 11.0, 6.0), 6.5), (2, (3.2, 3.4, 3.1), 7.0)])
def test_harness_list_matches_brute_force(seed, box, rcut):
    """the reference's "inputs" check (tests/test_python_repro_allegro.py:219-286) for the harness itself:
    the (tag_i, tag_j, r_ij) multiset of the full neighbour list within the cutoff equals an independent
    brute-force enumeration over periodic images (boxes smaller than the cutoff included)"""
    rng = np.random.default_rng(seed)
    n = 14
    cell = np.diag(box)
    pos = rng.random((n, 3)) * np.array(box)
    types = np.ones(n, dtype=np.int32)
    atoms = H.make_single_rank(types, pos, cell, [True] * 3, rcut)
    lst = H.build_full_list(atoms, rcut)
    got = []
    for i in range(atoms.nlocal):
        for j in lst.firstneigh(i):
            d = np.linalg.norm(atoms.x[j] - atoms.x[i])
            if d <= rcut:
                got.append((int(atoms.tag[i]), int(atoms.tag[j]), round(float(d), 9)))
    ref = []
    wrapped = atoms.x[:atoms.nlocal]
    kmax = [int(np.ceil(rcut / b)) + 1 for b in box]
    for i in range(n):
        for j in range(n):
            for a in range(-kmax[0], kmax[0] + 1):
                for b in range(-kmax[1], kmax[1] + 1):
                    for c in range(-kmax[2], kmax[2] + 1):
                        if i == j and a == b == c == 0:
                            continue
                        d = np.linalg.norm(wrapped[j] + np.array([a, b, c]) * np.array(box) - wrapped[i])
                        if d <= rcut:
                            ref.append((int(atoms.tag[i]), int(atoms.tag[j]), round(float(d), 9)))
    assert len(got) == len(ref) and sorted(got) == sorted(ref)
