import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import alg_path, load_golden
from pair_allegro_b200.pair import PairAllegroB200
for name, opts in (("Cu_r5", {}), ("Cu_r5", {"pipeline": "fused", "fused_batch": "2"}), ("aspirin_r5", {"pipeline": "fused"}), ("aspirin_r5", {})):
    atom, lst, z = load_golden(name)
    pair = PairAllegroB200(device=0, debug_mode=False)
    pair.coeff(["*", "*", alg_path(name)] + str(z["type_names"]).split(), atom.ntypes)
    for k, v in opts.items():
        pair.handle.set_option(k, v)
    pair.compute(atom, lst)
    print(name, opts, "max|dF| %.2e" % np.abs(atom.f - z["f"]).max())
