"""small runs of every new round-2 code path for compute-sanitizer (memcheck): default pipeline (device-built chunk plan,
persistent per-phase kernels), fused kernel (l_max 1, 2, 3), thread-per-atom edge build, halo exchange"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import alg_path, load_golden
from pair_allegro_b200 import capi
from pair_allegro_b200.pair import PairAllegroB200

def run(name, **opts):
    atom, lst, z = load_golden(name)
    pair = PairAllegroB200(device=0, debug_mode=False)
    pair.coeff(["*", "*", alg_path(name)] + str(z["type_names"]).split(), atom.ntypes)
    for k, v in opts.items():
        pair.handle.set_option(k, v)
    pair.compute(atom, lst)
    df = np.abs(atom.f - z["f"]).max()
    print(name, opts, "max|dF| %.2e" % df, "pipeline", pair.handle.stats("pipeline", 4))
    assert df < 1e-4
    return atom, lst, z, pair

for name in ("Cu_r5", "CuPd_r5", "Cu2AgO4_r5", "Cu_r15"):
    run(name, chunk_edges="4096")
for name in ("Cu_r5", "CuPd_r5", "Cu2AgO4_r5"):
    run(name, pipeline="fused", fused_batch="2")
# device entry, LayoutLeft view (thread-per-atom edge build), asynchronous call + halo
atom, lst, z, pair = run("CuPd_r5")
nl, ng = atom.nlocal, atom.nghost
maxn = int(lst.numneigh.max())
nb = np.zeros((nl, maxn), dtype=np.int32)
for i in range(nl):
    nb[i, :lst.numneigh[i]] = lst.firstneigh(i)
dev = torch.device("cuda:0")
d_x = torch.from_numpy(atom.x).to(dev); d_type = torch.from_numpy(atom.type).to(dev)
d_il = torch.arange(nl, dtype=torch.int32, device=dev); d_num = torch.from_numpy(lst.numneigh[:nl].copy()).to(dev)
d_nb = torch.from_numpy(np.ascontiguousarray(nb.T)).to(dev)
d_f = torch.zeros(nl + ng, 3, dtype=torch.float64, device=dev)
own = atom.owner[nl:].astype(np.int32)
comm = capi.Comm(0, 1, 0)
comm.set_plan(dict(recv_slices={0: (nl, nl + ng)}, send_index={0: own}, send_shift={0: atom.x[nl:] - atom.x[own]}))
st = torch.cuda.current_stream().cuda_stream
pair.handle.set_option("max_neighbors", str(maxn))
comm.forward(d_x.data_ptr(), st)
pair.handle.compute_device(nl, ng, d_x.data_ptr(), d_type.data_ptr(), d_il.data_ptr(), d_num.data_ptr(), d_nb.data_ptr(), 1, nl, d_f.data_ptr(), 0, want_scalars=False, stream=st)
comm.reverse(d_f.data_ptr(), st)
torch.cuda.synchronize()
from lmpshim import harness as H
assert np.abs(d_f[:nl].cpu().numpy() - H.reverse_comm_single_rank(atom, z["f"])).max() < 1e-4
print("device entry + halo OK")
