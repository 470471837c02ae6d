"""time alg_neigh_build (+ alg_neigh_check) on the C2 box and compare its rows with the torch builder:
   python tools/dev/neigh_time.py [ncell=63]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from lmpshim import harness as H  # noqa: E402
from lmpshim.nlist_torch import build_full_list_torch  # noqa: E402
from pair_allegro_b200 import capi  # noqa: E402

ncell = int(sys.argv[1]) if len(sys.argv) > 1 else 63
cfg = dict(bench.CONFIGS["c2"], ncell=ncell)
pos, types, cell = bench.config_box(cfg, 1, "weak")
rn = cfg["r_max"] + bench.SKIN
atoms = H.make_single_rank(types, pos, cell, [True] * 3, rn)
dev = torch.device("cuda:0")
nl, ng = atoms.nlocal, atoms.nghost
ref = build_full_list_torch(atoms.x, nl, rn, device=dev, want_host=False)
maxn = ref["maxn"]
d_x = torch.from_numpy(atoms.x).to(dev)
d_nb = torch.zeros(nl, maxn, dtype=torch.int32, device=dev)
d_num = torch.zeros(nl, dtype=torch.int32, device=dev)
lo, hi = atoms.x.min(0) - 1e-9, atoms.x.max(0) + 1e-9
nb_ = capi.NeighborBuilder(0)
mx = nb_.build(nl, ng, d_x.data_ptr(), lo, hi, rn, maxn, d_nb.data_ptr(), d_num.data_ptr())
torch.cuda.synchronize()
assert mx == maxn, (mx, maxn)
assert torch.equal(d_num, ref["numneigh"])
assert torch.equal(torch.sort(d_nb, 1).values, torch.sort(ref["nb2d"], 1).values)        # unused slots are 0 on both sides
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for want_max in (True, False):
    ts = []
    for it in range(6):
        e0.record()
        nb_.build(nl, ng, d_x.data_ptr(), lo, hi, rn, maxn, d_nb.data_ptr(), d_num.data_ptr(), want_max=want_max)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print("alg_neigh_build nlocal %d nghost %d pairs %d want_max=%s: %.3f ms (min of %s)" % (nl, ng, int(d_num.sum()), want_max, min(ts[1:]), ["%.2f" % t for t in ts]))
ts = []
for it in range(5):
    e0.record()
    r = nb_.needs_rebuild(nl + ng, d_x.data_ptr(), bench.SKIN)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print("alg_neigh_check: %.3f ms rebuild=%s" % (min(ts), r))
