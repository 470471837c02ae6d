"""Device-resident ghost halo: the C-ABI pack/unpack kernels plus
alg_compute_device reproduce what LAMMPS forward_comm / reverse_comm do around the pair style:
ghost x = owner x + image shift before the call, ghost f folded back onto the owners after it."""
import ctypes as C

import numpy as np
import pytest
import torch

from helpers import alg_path, load_golden
from test_gpu_parity import F_ATOL, make_pair

pytestmark = pytest.mark.gpu


def test_halo_pack_unpack_and_device_step(ensure_built):
    from lmpshim import harness as H
    from pair_allegro_b200 import capi
    lib = capi.load_library()
    name = "CuPd_r5"
    atom, lst, z = load_golden(name)
    nl, ng = atom.nlocal, atom.nghost
    ntot = nl + ng
    dev = torch.device("cuda:0")
    owner = atom.owner[nl:].astype(np.int32)
    shift = atom.x[nl:] - atom.x[owner]
    d_x = torch.from_numpy(atom.x).to(dev)
    d_x[nl:] = float("nan")                                   # ghosts must come from the halo
    d_idx = torch.from_numpy(owner).to(dev)
    d_shift = torch.from_numpy(np.ascontiguousarray(shift)).to(dev)
    st = torch.cuda.current_stream().cuda_stream
    # forward: pack straight into the ghost slice of x
    assert lib.alg_halo_pack(d_x.data_ptr(), d_idx.data_ptr(), ng, d_shift.data_ptr(), d_x[nl:].data_ptr(), st) == 0
    torch.cuda.synchronize()
    assert np.array_equal(d_x.cpu().numpy(), atom.x)          # bit-exact ghost positions
    # compute on device pointers
    pair = make_pair(name, z, atom)
    maxn = int(lst.numneigh.max())
    nb = np.zeros((nl, maxn), dtype=np.int32)
    for i in range(nl):
        nb[i, :lst.numneigh[i]] = lst.firstneigh(i)
    d_nb = torch.from_numpy(nb).to(dev)
    d_num = torch.from_numpy(lst.numneigh[:nl].copy()).to(dev)
    d_type = torch.from_numpy(atom.type).to(dev)
    d_ilist = torch.arange(nl, dtype=torch.int32, device=dev)
    d_f = torch.zeros(ntot, 3, dtype=torch.float64, device=dev)
    pair.handle.compute_device(nl, ng, d_x.data_ptr(), d_type.data_ptr(), d_ilist.data_ptr(), d_num.data_ptr(), d_nb.data_ptr(),
                               maxn, 1, d_f.data_ptr(), 0, want_scalars=True, stream=st)
    # reverse: ghost forces -> owners (several images of one atom: list entries repeat)
    assert lib.alg_halo_unpack_add(d_f.data_ptr(), d_idx.data_ptr(), ng, d_f[nl:].data_ptr(), st) == 0
    torch.cuda.synchronize()
    f_loc = d_f[:nl].cpu().numpy()
    ref = H.reverse_comm_single_rank(atom, z["f"])
    assert np.abs(f_loc - ref).max() < F_ATOL
    assert np.abs(f_loc.sum(0)).max() < 1e-4                  # momentum conservation after the reverse halo


def test_stats_and_timings(ensure_built):
    name = "CuPd_r5"
    atom, lst, z = load_golden(name)
    pair = make_pair(name, z, atom, profile="1", chunk_edges="4096", pipeline="tiled")
    pair.compute(atom, lst)
    h = pair.handle
    st = h.stats("step", 4)
    assert int(st[1]) == z["edge_index"].shape[1] and st[2] >= 2 and st[0] > 10      # edges, chunks, own launches
    kms = h.stats("kernel_ms", 6)
    kn = h.stats("kernel_launches", 6)
    assert kms[:5].min() > 0 and kn[0] >= st[2]               # one F0 launch per chunk (device-built plan: plus empty ones up to the capacity)
    t = h.timings()
    assert t.min() >= 0 and t[1] > 0


def test_comm_single_rank_self_images(ensure_built):
    """alg_comm_* with one rank: every ghost is a periodic image of a local atom, the exchange never leaves the device.
    forward reproduces the ghost positions bit-exactly; reverse folds the ghost forces onto their owners as a sorted
    segmented sum -- bit-identical run to run although several images of one atom are added (no fp64 atomics)."""
    from lmpshim import harness as H
    from pair_allegro_b200 import capi
    name = "Cu_r5"                                             # 4-atom cell, r_max 5: ~30 images per atom
    atom, lst, z = load_golden(name)
    nl, ng = atom.nlocal, atom.nghost
    dev = torch.device("cuda:0")
    owner = atom.owner[nl:].astype(np.int32)
    plan = dict(recv_slices={0: (nl, nl + ng)}, send_index={0: owner}, send_shift={0: atom.x[nl:] - atom.x[owner]})
    comm = capi.Comm(0, 1, 0)
    comm.set_plan(plan)
    st = torch.cuda.current_stream().cuda_stream
    d_x = torch.from_numpy(atom.x).to(dev)
    d_x[nl:] = float("nan")
    comm.forward(d_x.data_ptr(), st)
    torch.cuda.synchronize()
    assert np.array_equal(d_x.cpu().numpy(), atom.x)
    ref = H.reverse_comm_single_rank(atom, z["f"])
    outs = []
    for _ in range(3):
        d_f = torch.from_numpy(z["f"].copy()).to(dev)
        comm.reverse(d_f.data_ptr(), st)
        torch.cuda.synchronize()
        outs.append(d_f[:nl].cpu().numpy())
    assert np.abs(outs[0] - ref).max() < 1e-12
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])
    s = comm.stats()
    assert s[0] == 0 and int(s[2]) == ng and int(s[3]) <= nl   # nothing leaves the device; <= nlocal distinct owners
    assert np.allclose(comm.allreduce_sum([1.5, 2.5]), [1.5, 2.5])
    comm.close()
