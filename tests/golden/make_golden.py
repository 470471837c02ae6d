#!/usr/bin/env python
"""Generate the committed golden vectors under tests/golden/ (run in the build container,
where /root/reference is mounted; the GPU box only reads the outputs).

For each geometry fixture of the reference's own test-suite
(/root/reference/tests/conftest.py:55-62; frame 0 of /root/reference/tests/test_data/*.xyz)
it builds the LAMMPS-side state (ghost images + full neighbour list, skin 1.0 as in
/root/reference/tests/test_python_repro_allegro.py:100), runs the oracle restatement of
`pair_style allegro` (oracle/ref_pair.py -> TorchScript model of oracle/allegro_torch.py,
random-init weights) and stores inputs, the edge list and the outputs.  The known-answer edge
counts of SURVEY.md section 4 are asserted here.

  python tests/golden/make_golden.py
"""
import json
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import allegro_torch as AT  # noqa: E402
from lmpshim import harness as H  # noqa: E402
from oracle.ref_pair import RefPairAllegro  # noqa: E402
from pair_allegro_b200.export import export_alg  # noqa: E402

REF_DATA = "/root/reference/tests/test_data"
OUT = os.path.dirname(os.path.abspath(__file__))

# name, file, type names (alphabetical = LAMMPS type order in the reference tests), r_max,
# expected edges, l_max, layers
CASES = [
    ("CuPd_r5", "CuPd-cubic-big.xyz", ["Cu", "Pd"], 5.0, 10752, 2, 3),
    ("Cu_r5", "Cu-cubic.xyz", ["Cu"], 5.0, 168, 1, 2),
    ("Cu_r15", "Cu-cubic.xyz", ["Cu"], 15.0, 4816, 1, 1),
    ("Cu2AgO4_r5", "Cu2AgO4.xyz", ["Ag", "Cu", "O"], 5.0, 262, 3, 3),
    ("aspirin_r5", "aspirin.xyz", ["C", "H", "O"], 5.0, 306, 2, 2),
    ("aspirin_r15", "aspirin.xyz", ["C", "H", "O"], 15.0, 420, 1, 2),
]


def build_case(name, fname, type_names, r_max, expect_edges, lmax, nlayers, workdir):
    species, pos, cell, pbc = H.read_extxyz_frame(os.path.join(REF_DATA, fname), 0)
    if cell is None:  # non-periodic: 50 A box, centred (conftest.py:186-190)
        cell = np.eye(3) * 50.0
        pos = pos - pos.mean(axis=0) + 25.0
    skin = 1.0
    atoms = H.make_single_rank(species, pos, cell, pbc, r_max + skin, type_names)
    lst = H.build_full_list(atoms, r_max + skin)
    T = len(type_names)
    pc = None
    if name == "Cu2AgO4_r5":
        # exercise per_edge_type_cutoff (asymmetric allowed, cpp:303-328)
        pc = [[5.0, 4.5, 4.0], [4.5, 5.0, 4.2], [4.0, 4.2, 3.8]]
    cfg = AT.default_config(type_names=type_names, r_max=r_max, l_max=lmax, num_layers=nlayers,
                            per_edge_type_cutoff=pc, avg_num_neighbors=float(max(1.0, expect_edges / len(pos))),
                            per_type_energy_scales=[1.0 + 0.25 * t for t in range(T)],
                            per_type_energy_shifts=[-0.5 * t for t in range(T)],
                            seed=100 + len(name))
    pth = os.path.join(workdir, name + ".nequip.pth")
    AT.save_torchscript(cfg, pth)
    alg = os.path.join(OUT, name + ".alg")
    export_alg(pth, alg)
    pair = RefPairAllegro(debug_mode=False)
    pair.settings([])
    pair.coeff(["*", "*", pth] + type_names, atoms.ntypes)
    pair.init_style()
    small = lst.numneigh.sum() < 20000
    pair.compute(atoms, lst, loops=small)
    if small:  # the literal-loop and vectorised restatements must agree bit for bit
        e_loop = pair.last_input["edge_index"].numpy().copy()
        e_vec = pair.preprocess(atoms, lst)["edge_index"].numpy()
        assert np.array_equal(e_loop, e_vec)
    edges = pair.last_input["edge_index"].numpy()
    if pc is None:
        assert edges.shape[1] == expect_edges, (name, edges.shape[1], expect_edges)
    # fp64-parameter ground truth on the same inputs
    m64 = AT.build_model(cfg, torch.float64)
    o64 = m64(pair.last_input)
    out = pair.last_output
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        x=atoms.x, type=atoms.type, tag=atoms.tag, nlocal=atoms.nlocal, nghost=atoms.nghost, ntypes=atoms.ntypes,
        owner=atoms.owner, numneigh=lst.numneigh, neigh_flat=lst.neigh_flat, first=lst.first, ilist=lst.ilist,
        edge_index=edges, type_mapper=np.array(pair.type_mapper), cutoff_matrix=pair.cutoff_matrix,
        atomic_energy=out["atomic_energy"].numpy(), forces=out["forces"].numpy(), virial=out["virial"].numpy(),
        edge_energy=out["edge_energy"].numpy(),
        f=atoms.f, eng_vdwl=pair.eng_vdwl, eatom=pair.eatom, virial6=pair.virial,
        atomic_energy64=o64["atomic_energy"].numpy(), forces64=o64["forces"].numpy(), virial64=o64["virial"].numpy(),
        config=json.dumps(cfg), type_names=" ".join(type_names))
    print(f"{name}: atoms {atoms.nlocal}+{atoms.nghost} cand {int(lst.numneigh.sum())} edges {edges.shape[1]} "
          f"E {pair.eng_vdwl:.6f} |F|max {np.abs(out['forces'].numpy()).max():.4f} "
          f"f32-f64: dE {np.abs(out['atomic_energy'].numpy() - o64['atomic_energy'].numpy()).max():.2e} "
          f"dF {np.abs(out['forces'].numpy() - o64['forces'].numpy()).max():.2e}")


def main():
    torch.set_num_threads(8)
    with tempfile.TemporaryDirectory() as wd:
        for c in CASES:
            build_case(*c, wd)


if __name__ == "__main__":
    main()
