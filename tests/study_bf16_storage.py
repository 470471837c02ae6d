#!/usr/bin/env python
"""Study for DESIGN.md section 7 item 3 (not a test; run by hand): what would storing the inter-kernel
state of the tensor-core pipeline in bf16 cost in accuracy?  Uses the fp64 analytic oracle with a
quantisation hook on exactly the arrays the CUDA pipeline keeps in HBM (activation record, V^k, dV^k)
and reports max |dF| and max relative |dE_edge| against the unquantised run on the golden cases.
    python tests/study_bf16_storage.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import alg_path, load_golden  # noqa: E402
from oracle.analytic_numpy import AnalyticAllegro  # noqa: E402
from pair_allegro_b200.export import read_alg  # noqa: E402


def bf16(t):
    """round-to-nearest-even to 8 significant bits (bf16), returned as float64"""
    a = np.asarray(t, dtype=np.float32).view(np.uint32).astype(np.uint64)
    a = ((a + 0x7FFF + ((a >> 16) & 1)) >> 16) << 16
    return a.astype(np.uint32).view(np.float32).astype(np.float64).reshape(np.shape(t))


def forces(I, ei, ntot):
    f = np.zeros((ntot, 3))
    np.add.at(f, ei[0], I["g"])
    np.add.at(f, ei[1], -I["g"])
    return f


def main():
    print("%-12s %-10s %12s %14s" % ("case", "bf16 on", "max|dF| eV/A", "max rel dE_e"))
    for name in ("Cu_r5", "CuPd_r5", "aspirin_r5", "Cu2AgO4_r5"):
        atom, lst, z = load_golden(name)
        hdr, ten = read_alg(alg_path(name))
        ei = z["edge_index"]
        tm = z["type_mapper"]
        zi, zj = tm[atom.type[ei[0]] - 1], tm[atom.type[ei[1]] - 1]
        rvec = (atom.x[ei[1]] - atom.x[ei[0]]).astype(np.float32).astype(np.float64)
        A = AnalyticAllegro(hdr, ten)
        ref = A.run(rvec, ei[0], zi, zj, atom.nlocal)
        f0 = forces(ref, ei, len(atom.x))
        for label, sq in (("zd", {"zd": bf16}), ("v,dv", {"v": bf16, "dv": bf16}), ("zd,v,dv", {"zd": bf16, "v": bf16, "dv": bf16})):
            I = A.run(rvec, ei[0], zi, zj, atom.nlocal, store_quant=sq)
            df = np.abs(forces(I, ei, len(atom.x)) - f0).max()
            de = np.abs(I["e_edge"] - ref["e_edge"]).max() / max(np.abs(ref["e_edge"]).max(), 1e-30)
            print("%-12s %-10s %12.2e %14.2e" % (name, label, df, de))


if __name__ == "__main__":
    main()
