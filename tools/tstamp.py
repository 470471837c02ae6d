#!/usr/bin/env python
"""phase timing of one steady-state CTA of k_t_tc and k_bk_tc (clock64 marks, option debug=1)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lmpshim import harness as H
from pair_allegro_b200 import modelgen
from pair_allegro_b200.pair import PairAllegroB200
pos, types, cell = H.fcc_box(24)
atoms = H.make_single_rank(types, pos, cell, [True] * 3, 6.0)
lst = H.build_full_list(atoms, 6.0)
cfg = modelgen.default_config(type_names=["Ag"], r_max=5.0, l_max=1, num_layers=2, avg_num_neighbors=28.0, seed=2)
os.makedirs("/tmp/qb", exist_ok=True)
modelgen.random_alg(cfg, "/tmp/qb/m.alg")
pair = PairAllegroB200(device=0, debug_mode=False)
pair.coeff(["*", "*", "/tmp/qb/m.alg", "Ag"], 1)
pair.handle.set_option("debug", "1")
for _ in range(2):
    atoms.f[:] = 0
    pair.compute(atoms, lst)
ts = pair.handle.get_output("tstamp").reshape(5, 32)
t = ts[2]
names = {1: "geom+sync", 2: "stage rows", 3: "latent z1 (tp fwd, mma s, x rows, mma x)", 4: "hidden fwd (2 epi + 2 mma)",
         5: "x^n epilogue", 6: "mma readout", 7: "readout epi", 8: "mma dx", 9: "dx epi + du + E loop", 10: "bwd hidden (2 mma + 2 epi)", 11: "dIN (2 mma + 2 epi)",
         12: "tp backward (+segsum)", 13: "dy store", 14: "tc_end"}
order = [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14]
prev = t[0]
print("k_t_tc CTA 148: total %d cycles" % (t[14] - t[0]))
for i in order:
    print("  %-32s %7d" % (names[i], t[i] - prev))
    prev = t[i]

t = ts[3]
print("k_bk_tc CTA 148: total %d cycles" % (t[8] - t[0]))
for i, nm in enumerate(["geom + x rows + seg build", "stage dgamma rows", "phase 2 (2 mma, dw epi, dx epi)", "du + dm + stage gamma rows", "bwd hidden (2 mma + 2 epi)", "dIN (2 mma + 2 epi)", "tp backward (+segsum)", "dy store"]):
    print("  %-34s %7d" % (nm, t[i + 1] - t[i]))
