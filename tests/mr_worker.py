"""one rank of the N>1 GPU parity run (launched by tests/test_gpu_multirank.py / tools under torch.distributed.run):
brick decomposition -> alg_comm_forward (ghost x over NCCL) -> alg_compute_device -> alg_comm_reverse (ghost f to the
owners) ; every rank writes its owned forces / per-atom energies, rank 0 also the single-box result."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def device_inputs(atoms, lst, dev):
    nl = atoms.nlocal
    maxn = int(lst.numneigh[:nl].max())
    nb = np.zeros((nl, maxn), dtype=np.int32)
    cols = np.arange(len(lst.neigh_flat)) - np.repeat(lst.first[:nl], lst.numneigh[:nl])
    nb[np.repeat(np.arange(nl), lst.numneigh[:nl]), cols] = lst.neigh_flat
    return dict(nb=torch.from_numpy(nb).to(dev), num=torch.from_numpy(lst.numneigh[:nl].copy()).to(dev), maxn=maxn,
                type=torch.from_numpy(atoms.type).to(dev), ilist=torch.arange(nl, dtype=torch.int32, device=dev))


def evaluate(pair, comm, atoms, di, dev, repeat=2):
    """forward halo, force evaluation, reverse halo -- `repeat` times (bitwise determinism is checked by the caller)"""
    nl, ng = atoms.nlocal, atoms.nghost
    st = torch.cuda.current_stream().cuda_stream
    outs = []
    for _ in range(repeat):
        d_x = torch.from_numpy(atoms.x).to(dev)
        d_x[nl:] = float("nan")                               # ghost positions must come from the halo
        d_f = torch.zeros(nl + ng, 3, dtype=torch.float64, device=dev)
        d_e = torch.zeros(nl + ng, dtype=torch.float64, device=dev)
        comm.forward(d_x.data_ptr(), st)
        eng, vir = pair.handle.compute_device(nl, ng, d_x.data_ptr(), di["type"].data_ptr(), di["ilist"].data_ptr(), di["num"].data_ptr(),
                                              di["nb"].data_ptr(), di["maxn"], 1, d_f.data_ptr(), d_e.data_ptr(), want_scalars=True, stream=st)
        comm.reverse(d_f.data_ptr(), st)
        tot = comm.allreduce_sum(np.concatenate([[eng], vir]), st)
        torch.cuda.synchronize()
        outs.append(dict(f=d_f[:nl].cpu().numpy(), e=d_e[:nl].cpu().numpy(), eng=eng, tot=tot))
    return outs


def main():
    out_dir, ncell, lmax, nlayers = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("gloo")                           # only carries the NCCL id and the barrier; the halo is alg_comm_*
    from lmpshim import harness as H
    from pair_allegro_b200 import capi, modelgen
    from pair_allegro_b200.pair import PairAllegroB200
    ident = [capi.Comm.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ident, src=0)
    pos, types, cell = H.fcc_box(ncell, a=4.09, jitter=0.05, seed=4)
    types = (np.arange(len(pos)) % 2 + 1).astype(np.int32)
    alg = os.path.join(out_dir, "m.alg")
    if rank == 0:
        modelgen.random_alg(modelgen.default_config(type_names=["A", "B"], r_max=5.0, avg_num_neighbors=28.0, seed=9, l_max=lmax, num_layers=nlayers), alg)
    dist.barrier()
    rcomm = 6.0

    def make_pair():
        pair = PairAllegroB200(device=local, debug_mode=False)
        pair.coeff(["*", "*", alg, "A", "B"], 2)
        pair.init_style()
        return pair

    atoms, plan = H.decompose_rank(pos, types, cell, [True] * 3, world, rank, rcomm)
    lst = H.build_full_list(atoms, rcomm)
    comm = capi.Comm(local, world, rank, ident[0])
    comm.set_plan(plan)
    pair = make_pair()
    outs = evaluate(pair, comm, atoms, device_inputs(atoms, lst, dev), dev)
    assert np.array_equal(outs[0]["f"], outs[1]["f"]) and np.array_equal(outs[0]["e"], outs[1]["e"]), "N>1 run is not bitwise reproducible"
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), tag=atoms.tag[:atoms.nlocal], f=outs[0]["f"], e=outs[0]["e"], eng=outs[0]["eng"], tot=outs[0]["tot"],
             halo=comm.stats())
    comm.close()
    if rank == 0:                                             # the same box on one GPU (periodic self-images only)
        a1 = H.make_single_rank(types, pos, cell, [True] * 3, rcomm)
        l1 = H.build_full_list(a1, rcomm)
        own = a1.owner[a1.nlocal:].astype(np.int32)
        c1 = capi.Comm(local, 1, 0)
        c1.set_plan(dict(recv_slices={0: (a1.nlocal, a1.nlocal + a1.nghost)}, send_index={0: own}, send_shift={0: a1.x[a1.nlocal:] - a1.x[own]}))
        o1 = evaluate(make_pair(), c1, a1, device_inputs(a1, l1, dev), dev, repeat=1)[0]
        np.savez(os.path.join(out_dir, "single.npz"), f=o1["f"], e=o1["e"], eng=o1["eng"], tot=o1["tot"])
        c1.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
