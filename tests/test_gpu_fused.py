"""The fused persistent kernel (centre-aligned tiles, all phases of a tile in one CTA; option pipeline=fused, the
default whenever no atom has more than 128 neighbours) against the oracle goldens and against the chunked
edge-tile pipeline (pipeline=tiled), plus the asynchronous device entry."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN_CASES
from helpers import alg_path, load_golden
from test_gpu_parity import E_ATOL, E_RTOL, F_ATOL, V_RTOL, check_outputs, make_pair

pytestmark = pytest.mark.gpu

FUSED_CASES = [c for c in GOLDEN_CASES if c != "Cu_r15"]     # Cu_r15: 1204 neighbours per atom -> chunked pipeline


@pytest.mark.parametrize("name", FUSED_CASES)
def test_fused_golden_parity(name, ensure_built):
    atom, lst, z = load_golden(name)
    pair = make_pair(name, z, atom, keep_edges="1", pipeline="fused")
    pair.compute(atom, lst)
    assert pair.handle.stats("pipeline", 3)[0] == 1
    e = pair.handle.get_edges()
    assert np.array_equal(e, z["edge_index"])
    check_outputs(pair, atom, z)
    st = pair.handle.stats("step", 4)
    assert int(st[1]) == z["edge_index"].shape[1] and st[3] >= 1


def test_long_rows(ensure_built):
    """an atom with more neighbours than a fused batch holds (1204 > 8 * 128): the default (chunked, device-built plan)
    handles it; pipeline=fused reports the limit; with a larger batch the fused kernel takes it too"""
    from pair_allegro_b200 import capi
    name = "Cu_r15"
    atom, lst, z = load_golden(name)
    pair = make_pair(name, z, atom)
    pair.compute(atom, lst)
    assert list(pair.handle.stats("pipeline", 4)[[0, 3]]) == [0, 1]
    check_outputs(pair, atom, z)
    pair2 = make_pair(name, z, atom, pipeline="fused")
    with pytest.raises(capi.AllegroError):
        pair2.compute(atom, lst)
    atom.f[:] = 0
    pair3 = make_pair(name, z, atom, pipeline="fused", fused_batch="16")
    pair3.compute(atom, lst)
    assert pair3.handle.stats("pipeline", 4)[0] == 1
    check_outputs(pair3, atom, z)


def test_chunk_plan_device_equals_host(ensure_built):
    """the device-built chunk plan (no host synchronisation) against the host-built plan of round 1: same chunks up to the
    alignment rule, results equal to fp32 round-off; a plan that does not fit the buffers falls back transparently"""
    name = "CuPd_r5"
    atom, lst, z = load_golden(name)
    out = {}
    for plan in ("device", "host"):
        atom.f[:] = 0
        pair = make_pair(name, z, atom, chunk_plan=plan, chunk_edges="4096")
        pair.compute(atom, lst)
        check_outputs(pair, atom, z)
        assert pair.handle.stats("pipeline", 4)[3] == (1 if plan == "device" else 0)
        out[plan] = (atom.f.copy(), pair.eatom.copy())
    assert np.abs(out["device"][0] - out["host"][0]).max() < 2e-5
    # isolated atoms: fewer than 4 neighbours per atom on average -> the centre capacity of the device plan can be exceeded;
    # here it is not (3 atoms), but the path with zero edges must work
    class A: pass
    a = A(); a.x = np.array([[0.0, 0, 0], [20.0, 0, 0], [0, 20.0, 0]]); a.type = np.array([1, 2, 1], dtype=np.int32); a.tag = np.arange(1, 4)
    a.nlocal, a.nghost, a.ntypes, a.f = 3, 0, atom.ntypes, np.zeros((3, 3))
    l = A(); l.inum, l.gnum = 3, 0; l.ilist = np.arange(3, dtype=np.int32); l.numneigh = np.array([2, 2, 2], dtype=np.int32)
    l.neigh_flat = np.array([1, 2, 0, 2, 0, 1], dtype=np.int32); l.first = np.array([0, 2, 4], dtype=np.int64)
    pair = make_pair(name, z, a)
    pair.compute(a, l)
    assert np.abs(a.f).max() == 0.0


def _fcc(ncell, seed=7):
    from lmpshim import harness as H
    pos, types, cell = H.fcc_box(ncell, a=4.09, jitter=0.08, seed=seed)
    atoms = H.make_single_rank(types, pos, cell, [True] * 3, 6.0)
    return atoms, H.build_full_list(atoms, 6.0)


@pytest.mark.parametrize("lmax,nlayers", [(1, 1), (1, 2), (1, 3), (2, 2), (2, 3), (3, 3)])
def test_fused_equals_tiled(lmax, nlayers, ensure_built, tmp_path):
    """same kernels bodies, different tiling: agreement to fp32 round-off on a 2048-atom box (many tiles per CTA slot)"""
    from pair_allegro_b200 import modelgen
    from pair_allegro_b200.pair import PairAllegroB200
    atoms, lst = _fcc(8)
    alg = str(tmp_path / "m.alg")
    modelgen.random_alg(modelgen.default_config(type_names=["Ag"], r_max=5.0, avg_num_neighbors=26.0, seed=11, l_max=lmax, num_layers=nlayers), alg)
    out, stats = {}, {}
    for mode in ("fused", "tiled"):
        pair = PairAllegroB200(device=0, debug_mode=False)
        pair.coeff(["*", "*", alg, "Ag"], 1)
        pair.init_style()
        pair.handle.set_option("pipeline", mode)
        runs = []
        for _ in range(2):
            atoms.f[:] = 0
            pair.compute(atoms, lst)
            runs.append((atoms.f.copy(), pair.eatom[:atoms.nlocal].copy(), pair.virial.copy(), pair.eng_vdwl))
        assert np.array_equal(runs[0][0], runs[1][0]) and np.array_equal(runs[0][1], runs[1][1]) and runs[0][3] == runs[1][3]
        out[mode] = runs[0]
        stats[mode] = pair.handle.stats("step", 4)
        assert pair.handle.stats("pipeline", 3)[0] == (1 if mode == "fused" else 0)
    # batch plan: centre-aligned batches of edge-aligned tiles waste at most one tile per batch
    E, nbatch, ntile = (int(v) for v in stats["fused"][1:4])
    assert (E + 127) // 128 <= ntile <= (E + 127) // 128 + nbatch and 1 <= nbatch <= ntile
    assert np.abs(out["fused"][0] - out["tiled"][0]).max() < 2e-5
    np.testing.assert_allclose(out["fused"][1], out["tiled"][1], rtol=2e-6, atol=2e-6)
    assert np.abs(out["fused"][2] - out["tiled"][2]).max() < 1e-5 * max(1.0, np.abs(out["tiled"][2]).max())


@pytest.mark.parametrize("pipeline", ["auto", "fused"])
def test_device_entry_is_asynchronous(pipeline, ensure_built, tmp_path):
    """alg_compute_device(eng=NULL, virial=NULL) must return before the stream has drained (no hidden host
    synchronisation, cf. the blocking nedges copy of pair_nequip_allegro_kokkos.cpp:203-206), and a later synchronous
    call must see the same forces"""
    from pair_allegro_b200 import modelgen
    from pair_allegro_b200.pair import PairAllegroB200
    atoms, lst = _fcc(20)                      # 32000 atoms: tens of milliseconds of device work
    alg = str(tmp_path / "m.alg")
    modelgen.random_alg(modelgen.default_config(type_names=["Ag"], r_max=5.0, avg_num_neighbors=26.0, seed=3, l_max=1, num_layers=2), alg)
    pair = PairAllegroB200(device=0, debug_mode=False)
    pair.coeff(["*", "*", alg, "Ag"], 1)
    pair.init_style()
    nl, ntot = atoms.nlocal, atoms.nlocal + atoms.nghost
    maxn = int(lst.numneigh.max())
    nb = np.zeros((nl, maxn), dtype=np.int32)
    for i in range(nl):
        nb[i, :lst.numneigh[i]] = lst.firstneigh(i)
    dev = torch.device("cuda:0")
    d_x = torch.from_numpy(atoms.x).to(dev); d_type = torch.from_numpy(atoms.type).to(dev)
    d_ilist = torch.arange(nl, dtype=torch.int32, device=dev); d_num = torch.from_numpy(lst.numneigh[:nl].copy()).to(dev)
    d_nb = torch.from_numpy(nb).to(dev)
    h = pair.handle
    h.set_option("pipeline", pipeline)
    h.set_option("max_neighbors", str(maxn))
    stream = torch.cuda.Stream()
    args = (nl, atoms.nghost, d_x.data_ptr(), d_type.data_ptr(), d_ilist.data_ptr(), d_num.data_ptr(), d_nb.data_ptr(), maxn, 1)
    d_f0 = torch.zeros(ntot, 3, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    eng, _ = h.compute_device(*args, d_f0.data_ptr(), 0, want_scalars=True, stream=stream.cuda_stream)      # synchronous reference (and warm-up)
    torch.cuda.synchronize()
    not_ready = 0
    for _ in range(3):
        d_f = torch.zeros(ntot, 3, dtype=torch.float64, device=dev)
        torch.cuda.synchronize()
        h.compute_device(*args, d_f.data_ptr(), 0, want_scalars=False, stream=stream.cuda_stream)
        not_ready += 0 if stream.query() else 1            # cudaStreamQuery == cudaErrorNotReady
        stream.synchronize()
        assert torch.equal(d_f, d_f0)
    assert not_ready == 3, "alg_compute_device(eng=NULL) blocked until the device finished"
    assert int(h.stats("step", 4)[1]) > 0                    # the deferred verdict of the last asynchronous step
