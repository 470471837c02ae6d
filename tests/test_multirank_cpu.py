"""world_size-2 (and 4) gloo tests of the N>1 path on CPU: brick decomposition + halo plan +
forward ghost-x / reverse ghost-f exchange reproduce the single-rank forces and energies
(the reference's own multi-rank check: tests/test_python_repro_allegro.py:44-47 runs 1/2/4 ranks
and expects identical forces).  The per-rank force evaluation uses the CPU oracle; the transport is
torch.distributed (gloo) point-to-point exactly as bench.py's Halo class does it over NCCL."""
import os
import sys
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from oracle import allegro_torch as AT
    from lmpshim import harness as H
    from oracle.ref_pair import RefPairAllegro
    pos, types, cell = H.fcc_box(6, a=4.09, jitter=0.05, seed=4)
    types = (np.arange(len(pos)) % 2 + 1).astype(np.int32)
    rcomm = 6.0
    atoms, plan = H.decompose_rank(pos, types, cell, [True] * 3, world, rank, rcomm)
    nl = atoms.nlocal
    # ---- forward comm: owners pack x[send]+shift, receivers write their ghost slices
    x = atoms.x.copy()
    x[nl:] = np.nan
    ops, keep = [], []
    for s, idx in plan["send_index"].items():
        buf = torch.from_numpy(x[idx] + plan["send_shift"][s])
        if s == rank:
            a, b = plan["recv_slices"][s]
            x[a:b] = buf.numpy()
        else:
            keep.append(buf)
            ops.append(dist.P2POp(dist.isend, buf, s))
    recv = {}
    for s, (a, b) in plan["recv_slices"].items():
        if s != rank:
            recv[s] = torch.empty(b - a, 3, dtype=torch.float64)
            ops.append(dist.P2POp(dist.irecv, recv[s], s))
    for w in (dist.batch_isend_irecv(ops) if ops else []):
        w.wait()
    for s, t in recv.items():
        a, b = plan["recv_slices"][s]
        x[a:b] = t.numpy()
    assert np.abs(x - atoms.x).max() < 1e-12, "forward halo does not reproduce ghost positions"
    # ---- per-rank force evaluation (oracle), newton on: forces on locals AND ghosts
    lst = H.build_full_list(atoms, rcomm)
    pair = RefPairAllegro()
    pair.coeff(["*", "*", os.path.join(tmp, "m.nequip.pth"), "A", "B"], 2)
    pair.compute(atoms, lst)
    f = atoms.f
    # ---- reverse comm: ghost slices go back to the owners and are accumulated
    ops, keep, rbuf = [], [], {}
    for s, (a, b) in plan["recv_slices"].items():
        if s != rank:
            t = torch.from_numpy(f[a:b].copy())
            keep.append(t)
            ops.append(dist.P2POp(dist.isend, t, s))
    for s, idx in plan["send_index"].items():
        if s != rank:
            rbuf[s] = torch.empty(len(idx), 3, dtype=torch.float64)
            ops.append(dist.P2POp(dist.irecv, rbuf[s], s))
    for w in (dist.batch_isend_irecv(ops) if ops else []):
        w.wait()
    ftot = f[:nl].copy()
    for s, idx in plan["send_index"].items():
        src = rbuf[s].numpy() if s != rank else f[plan["recv_slices"][s][0]:plan["recv_slices"][s][1]]
        np.add.at(ftot, idx, src)
    np.savez(os.path.join(tmp, "rank%d.npz" % rank), tag=atoms.tag[:nl], f=ftot, e=pair.eatom[:nl], eng=pair.eng_vdwl)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_halo_exchange_reproduces_single_rank(world):
    sys.path.insert(0, ROOT)
    from oracle import allegro_torch as AT
    from lmpshim import harness as H
    from oracle.ref_pair import RefPairAllegro
    with tempfile.TemporaryDirectory() as tmp:
        cfg = AT.default_config(type_names=["A", "B"], r_max=5.0, l_max=1, num_layers=2, avg_num_neighbors=28.0, seed=9)
        AT.save_torchscript(cfg, os.path.join(tmp, "m.nequip.pth"))
        # single-rank truth
        pos, types, cell = H.fcc_box(6, a=4.09, jitter=0.05, seed=4)
        types = (np.arange(len(pos)) % 2 + 1).astype(np.int32)
        atoms = H.make_single_rank(types, pos, cell, [True] * 3, 6.0)
        lst = H.build_full_list(atoms, 6.0)
        pair = RefPairAllegro()
        pair.coeff(["*", "*", os.path.join(tmp, "m.nequip.pth"), "A", "B"], 2)
        pair.compute(atoms, lst)
        f1 = H.reverse_comm_single_rank(atoms, atoms.f)
        e1 = pair.eatom[:atoms.nlocal]
        port = 29600 + world
        mp.spawn(_worker, args=(world, port, tmp), nprocs=world, join=True)
        fN = np.zeros_like(f1)
        eN = np.zeros_like(e1)
        seen = np.zeros(len(f1), dtype=int)
        eng = 0.0
        for r in range(world):
            z = np.load(os.path.join(tmp, "rank%d.npz" % r))
            fN[z["tag"] - 1] = z["f"]
            eN[z["tag"] - 1] = z["e"]
            seen[z["tag"] - 1] += 1
            eng += float(z["eng"])
    assert np.all(seen == 1)
    assert np.abs(fN - f1).max() < 2e-5          # fp32 model, different edge order per rank
    np.testing.assert_allclose(eN, e1, rtol=1e-5, atol=1e-5)
    assert abs(eng - pair.eng_vdwl) < 1e-4
