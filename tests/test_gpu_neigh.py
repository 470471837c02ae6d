"""alg_neigh_* (binned FULL neighbour list on the device, the caller-side step LAMMPS' Neighbor performs before the pair
style) against the numpy cell-list builder and brute force; determinism; the Verlet-skin check; and the whole chain
neighbour build -> alg_compute_device against the golden forces."""
import numpy as np
import pytest
import torch

from helpers import alg_path, load_golden
from test_gpu_parity import F_ATOL, make_pair

pytestmark = pytest.mark.gpu


def _build(nb_, atoms, rn, max_neigh, layout_left=False):
    dev = torch.device("cuda:0")
    nl, ng = atoms.nlocal, atoms.nghost
    d_x = torch.from_numpy(atoms.x).to(dev)
    shape = (max_neigh, nl) if layout_left else (nl, max_neigh)
    d_nb = torch.full(shape, -1, dtype=torch.int32, device=dev)
    d_num = torch.zeros(nl, dtype=torch.int32, device=dev)
    lo, hi = atoms.x.min(0) - 1e-9, atoms.x.max(0) + 1e-9
    si, sj = (1, nl) if layout_left else (max_neigh, 1)
    mx = nb_.build(nl, ng, d_x.data_ptr(), lo, hi, rn, max_neigh, d_nb.data_ptr(), d_num.data_ptr(), stride_i=si, stride_jj=sj)
    torch.cuda.synchronize()
    nb = d_nb.cpu().numpy()
    return (nb.T if layout_left else nb), d_num.cpu().numpy(), mx, d_x, d_nb, d_num


@pytest.mark.parametrize("box", ["fcc", "water"])
def test_device_list_equals_numpy_builder(box, ensure_built):
    from lmpshim import harness as H
    from pair_allegro_b200 import capi
    (pos, types, cell), rn = (H.fcc_box(7, jitter=0.08, seed=1), 6.0) if box == "fcc" else (H.water_like_box(500, seed=2), 7.0)
    atoms = H.make_single_rank(types, pos, cell, [True] * 3, rn)
    ref = H.build_full_list(atoms, rn)
    nb_ = capi.NeighborBuilder(0)
    maxn = int(ref.numneigh.max())
    for layout_left in (False, True):
        nb, num, mx, *_ = _build(nb_, atoms, rn, maxn + 3, layout_left)
        assert mx == maxn and np.array_equal(num, ref.numneigh[:atoms.nlocal])
        for i in range(0, atoms.nlocal, 7):
            assert np.array_equal(np.sort(nb[i, :num[i]]), ref.neigh_flat[ref.first[i]:ref.first[i] + num[i]])     # ref rows are ascending
    nb2, num2, *_ = _build(nb_, atoms, rn, maxn + 3)
    nb3, num3, *_ = _build(nb_, atoms, rn, maxn + 3)
    assert np.array_equal(nb2, nb3) and np.array_equal(num2, num3)             # same order every time
    # a view that is too small is reported, not silently truncated
    with pytest.raises(capi.AllegroError):
        _build(nb_, atoms, rn, maxn - 1)


def test_verlet_skin_check(ensure_built):
    from lmpshim import harness as H
    from pair_allegro_b200 import capi
    pos, types, cell = H.fcc_box(5, jitter=0.05, seed=3)
    atoms = H.make_single_rank(types, pos, cell, [True] * 3, 6.0)
    nb_ = capi.NeighborBuilder(0)
    _, _, _, d_x, _, _ = _build(nb_, atoms, 6.0, 96)
    ntot = atoms.nlocal + atoms.nghost
    assert not nb_.needs_rebuild(ntot, d_x.data_ptr(), 1.0)
    d_x[17, 0] += 0.49
    assert not nb_.needs_rebuild(ntot, d_x.data_ptr(), 1.0)                     # moved less than skin / 2
    d_x[17, 1] += 0.2
    assert nb_.needs_rebuild(ntot, d_x.data_ptr(), 1.0)                         # sqrt(0.49^2 + 0.2^2) > 0.5
    assert nb_.needs_rebuild(ntot - 1, d_x.data_ptr(), 1.0)                     # atom count changed


def test_device_list_feeds_the_force_evaluation(ensure_built):
    """positions on the device -> alg_neigh_build -> alg_compute_device: the golden forces without any host-built list"""
    from pair_allegro_b200 import capi
    name = "CuPd_r5"
    atom, lst, z = load_golden(name)
    pair = make_pair(name, z, atom)
    nb_ = capi.NeighborBuilder(0)
    rn = 6.0
    nb, num, mx, d_x, d_nb, d_num = _build(nb_, atom, rn, 128)
    dev = torch.device("cuda:0")
    nl, ng = atom.nlocal, atom.nghost
    d_type = torch.from_numpy(atom.type).to(dev)
    d_il = torch.arange(nl, dtype=torch.int32, device=dev)
    d_f = torch.zeros(nl + ng, 3, dtype=torch.float64, device=dev)
    pair.handle.set_option("max_neighbors", "128")
    eng, vir = pair.handle.compute_device(nl, ng, d_x.data_ptr(), d_type.data_ptr(), d_il.data_ptr(), d_num.data_ptr(), d_nb.data_ptr(), 128, 1,
                                          d_f.data_ptr(), 0, want_scalars=True)
    torch.cuda.synchronize()
    assert np.abs(d_f.cpu().numpy() - z["f"]).max() < F_ATOL
    assert abs(eng - float(z["eng_vdwl"])) < 1e-4
