"""the torch neighbour-list builder (lmpshim/nlist_torch.py, used by bench.py and the full-size tests) against the numpy
cell-list builder and brute force: same rows as sets, same counts, 2-D view consistent with the CSR"""
import numpy as np

from lmpshim import harness as H
from lmpshim.nlist_torch import as_neighlist, build_full_list_torch


def _rows(lst, nl):
    return [np.sort(lst.neigh_flat[lst.first[i]:lst.first[i] + lst.numneigh[i]]) for i in range(nl)]


def test_matches_numpy_builder_fcc_and_water():
    for pos, types, cell, rn in (H.fcc_box(5, jitter=0.08, seed=1) + (6.0,), H.water_like_box(300, seed=2) + (7.0,)):
        atoms = H.make_single_rank(types, pos, cell, [True] * 3, rn)
        ref = H.build_full_list(atoms, rn)
        res = build_full_list_torch(atoms.x, atoms.nlocal, rn, device="cpu", chunk=97)
        lst = as_neighlist(atoms, res)
        assert np.array_equal(lst.numneigh, ref.numneigh)
        for a, b in zip(_rows(lst, atoms.nlocal), _rows(ref, atoms.nlocal)):
            assert np.array_equal(a, b)
        nb = res["nb2d"].numpy()
        nn = res["numneigh"].numpy()
        for i in (0, atoms.nlocal // 2, atoms.nlocal - 1):
            assert np.array_equal(nb[i, :nn[i]], lst.neigh_flat[lst.first[i]:lst.first[i] + nn[i]])
        assert res["candidates"] == int(ref.numneigh.sum())


def test_brute_force_small_cluster():
    rng = np.random.default_rng(3)
    x = rng.uniform(0, 9, (60, 3))
    res = build_full_list_torch(x, 40, 3.0, device="cpu")
    d2 = ((x[:40, None, :] - x[None, :, :]) ** 2).sum(-1)
    for i in range(40):
        want = np.nonzero((d2[i] <= 9.0) & (np.arange(60) != i))[0]
        got = np.sort(res["neigh_flat_h"][res["first_h"][i]:res["first_h"][i] + res["numneigh_h"][i]])
        assert np.array_equal(got, want)
