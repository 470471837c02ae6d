"""shared helpers of the parity tests"""
import json
import os

import numpy as np

from conftest import GOLDEN


class _NS:
    pass


def load_golden(name):
    """golden case -> (atom, list, npz) with the LAMMPS stand-in fields (lmpshim/harness.py)"""
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    atom = _NS()
    atom.x = z["x"].copy()
    atom.type = z["type"].astype(np.int32)
    atom.tag = z["tag"]
    atom.nlocal = int(z["nlocal"])
    atom.nghost = int(z["nghost"])
    atom.ntypes = int(z["ntypes"])
    atom.f = np.zeros_like(atom.x)
    atom.owner = z["owner"]
    lst = _NS()
    lst.inum = atom.nlocal
    lst.gnum = atom.nghost
    lst.ilist = z["ilist"].astype(np.int32)
    lst.numneigh = z["numneigh"].astype(np.int32)
    lst.neigh_flat = z["neigh_flat"].astype(np.int32)
    lst.first = z["first"].astype(np.int64)
    lst.firstneigh = lambda i: lst.neigh_flat[lst.first[i]:lst.first[i] + lst.numneigh[i]]
    return atom, lst, z


def golden_config(z):
    return json.loads(str(z["config"]))


def alg_path(name):
    return os.path.join(GOLDEN, name + ".alg")
