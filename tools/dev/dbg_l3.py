import sys, os, numpy as np, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from lmpshim import harness as H
from pair_allegro_b200 import modelgen
from pair_allegro_b200.pair import PairAllegroB200
L = int(sys.argv[1]) if len(sys.argv) > 1 else 3
nl_ = int(sys.argv[2]) if len(sys.argv) > 2 else 3
pos, types, cell = H.multi_species_box(700, fractions=(3, 1, 4), density=0.09, seed=5)
cfg = modelgen.default_config(type_names=["Li", "P", "O"], r_max=5.5, l_max=L, num_layers=nl_, avg_num_neighbors=60.0, seed=5)
d = tempfile.mkdtemp(); alg = d + "/m.alg"
modelgen.random_alg(cfg, alg)
atoms = H.make_single_rank(types, pos, cell, [True] * 3, 6.5)
lst = H.build_full_list(atoms, 6.5)
res = {}
for gemm, pipe in (("ffma", "tiled"), ("tc", "tiled"), ("tc", "fused")):
    atoms.f[:] = 0
    p = PairAllegroB200(device=0, debug_mode=False)
    p.coeff(["*", "*", alg, "Li", "P", "O"], 3)
    p.handle.set_option("gemm", gemm); p.handle.set_option("pipeline", pipe)
    p.compute(atoms, lst)
    res[(gemm, pipe)] = (atoms.f.copy(), p.eatom[:atoms.nlocal].copy(), p.eng_vdwl, p.virial.copy())
    print(gemm, pipe, "eng", p.eng_vdwl, "stats", p.handle.stats("step", 4), p.handle.stats("pipeline", 3))
ref = res[("ffma", "tiled")]
for k, v in res.items():
    print(k, "max|dF|", np.abs(v[0] - ref[0]).max(), "max|dE|", np.abs(v[1] - ref[1]).max(), "dvir", np.abs(v[3] - ref[3]).max(), "nan:", np.isnan(v[0]).sum())
