#!/usr/bin/env python
"""quick single-GPU timing of the force evaluation on an FCC box (development aid)"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lmpshim import harness as H  # noqa: E402
from pair_allegro_b200 import modelgen  # noqa: E402
from pair_allegro_b200.pair import PairAllegroB200  # noqa: E402


def main():
    ncell = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    nl = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    chunk = sys.argv[4] if len(sys.argv) > 4 else "1048576"
    gemm = sys.argv[5] if len(sys.argv) > 5 else "ffma"
    prec = sys.argv[6] if len(sys.argv) > 6 else "strict"
    pipeline = sys.argv[7] if len(sys.argv) > 7 else "auto"
    pos, types, cell = H.fcc_box(ncell)
    t0 = time.time()
    atoms = H.make_single_rank(types, pos, cell, [True] * 3, 6.0)
    lst = H.build_full_list(atoms, 6.0)
    print("atoms %d ghosts %d cand %d  (harness %.1fs)" % (atoms.nlocal, atoms.nghost, lst.numneigh.sum(), time.time() - t0))
    cfg = modelgen.default_config(type_names=["Ag"], r_max=5.0, l_max=L, num_layers=nl, avg_num_neighbors=28.0, seed=2)
    if os.environ.get("ALG_WIDTHS"):                      # S,U,H,depth,R (other than 64,32,64,2,32 -> width-generic pipeline)
        S, U, Hh, D, R = (int(v) for v in os.environ["ALG_WIDTHS"].split(","))
        cfg.update(num_scalar_features=S, num_tensor_features=U, mlp_width=Hh, mlp_depth=D, readout_width=R)
    os.makedirs("/tmp/qb", exist_ok=True)
    modelgen.random_alg(cfg, "/tmp/qb/m.alg")
    pair = PairAllegroB200(device=0, debug_mode=False)
    pair.coeff(["*", "*", "/tmp/qb/m.alg", "Ag"], 1)
    pair.handle.set_option("chunk_edges", chunk)
    pair.handle.set_option("gemm", gemm)
    if gemm == "tc":
        pair.handle.set_option("precision", prec)
    pair.handle.set_option("profile", "1")
    pair.handle.set_option("pipeline", pipeline)
    if os.environ.get("ALG_CHUNK_PLAN"):
        pair.handle.set_option("chunk_plan", os.environ["ALG_CHUNK_PLAN"])
    if os.environ.get("ALG_FUSED_BATCH"):
        pair.handle.set_option("fused_batch", os.environ["ALG_FUSED_BATCH"])
    for it in range(4):
        atoms.f[:] = 0
        t0 = time.time()
        pair.compute(atoms, lst)
        dt = time.time() - t0
        tm = pair.handle.timings()
        print("iter %d: wall %.1f ms  device: edges %.2f ms, network %.2f ms, finalize %.2f ms  -> %.3f Matom-steps/s (device)  eng %.6f" %
              (it, dt * 1e3, tm[0], tm[1], tm[2], atoms.nlocal / (tm.sum() * 1e-3) / 1e6, pair.eng_vdwl))
    print("kernel ms:", dict(zip(["F0", "FK", "T", "BK", "B0", "fixup", "fused"], np.round(pair.handle.stats("kernel_ms", 7), 2))),
          "pipeline:", pair.handle.stats("pipeline", 3), "step:", pair.handle.stats("step", 4))


if __name__ == "__main__":
    main()
