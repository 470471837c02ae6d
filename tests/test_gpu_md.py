"""The reference's own test pattern (tests/test_python_repro_allegro.py) on the CUDA path:
 * C1 of BASELINE.json: 512-atom CuPd box, test-yaml model (l_max 2, 3 layers), 100 NVE steps with the
   neighbour list rebuilt every step (skin 1.0); at regular steps the forces / per-atom energies /
   virial the integrator is fed are compared with the reference path at the SAME positions
   (test_python_repro_allegro.py:302-355), and pe == sum(pe/atom) (:321);
 * invariance of the result to the number of ranks (1 / 2 / 4 bricks, :44-47, 71-77), ghosts' forces
   folded back by the reverse halo."""
import numpy as np
import pytest

from helpers import load_golden
from pair_allegro_b200 import modelgen

pytestmark = pytest.mark.gpu

FTM2V = 1.0 / 1.0364269e-4     # LAMMPS `units metal`: (eV/A)/(g/mol) -> A/ps^2
MVV2E = 1.0364269e-4           # g/mol (A/ps)^2 -> eV


def _c1_box():
    atom, _, z = load_golden("CuPd_r5")
    nl = atom.nlocal
    sh = np.abs(atom.x[nl:] - atom.x[atom.owner[nl:]])
    box = float(sh[sh > 1.0].min())                       # cubic cell edge recovered from the ghost images
    pos = np.concatenate([atom.x[:nl], atom.x[:nl] + np.array([box, 0.0, 0.0])])
    types = np.concatenate([atom.type[:nl], atom.type[:nl]])
    return pos, types, np.diag([2 * box, box, box])


def test_c1_nve_100_steps(ensure_built, tmp_path):
    from lmpshim import harness as H
    from oracle import allegro_torch as AT
    from oracle.ref_pair import RefPairAllegro
    from pair_allegro_b200.pair import PairAllegroB200
    from test_gpu_parity import E_ATOL, E_RTOL, F_ATOL, V_RTOL
    pos, types, cell = _c1_box()
    assert len(pos) == 512
    names = ["Cu", "Pd"]
    cfg = modelgen.default_config(type_names=names, seed=1)                 # test_repro_allegro.yaml hyper-parameters
    alg, pth = str(tmp_path / "c1.alg"), str(tmp_path / "c1.nequip.pth")
    modelgen.random_alg(cfg, alg)
    AT.save_torchscript_from_alg(alg, pth)
    ours = PairAllegroB200(device=0, debug_mode=False)
    ours.coeff(["*", "*", alg] + names, 2)
    ours.init_style()
    ref = RefPairAllegro()
    ref.coeff(["*", "*", pth] + names, 2)
    mass = np.where(types == 1, 63.546, 106.42)[:, None]
    dt, nsteps, skin = 0.001, 100, 1.0
    x = pos.copy()
    v = np.zeros_like(x)

    def force(xc, check):
        atoms = H.make_single_rank(types, xc, cell, [True] * 3, 5.0 + skin)
        lst = H.build_full_list(atoms, 5.0 + skin)
        ours.compute(atoms, lst, eflag=1, vflag=1)
        f = H.reverse_comm_single_rank(atoms, atoms.f)
        pe = ours.eng_vdwl
        n = atoms.nlocal
        assert abs(pe - ours.eatom[:n].sum()) < 1e-8 * max(1.0, abs(pe))     # pe == sum pe/atom
        if check:
            f_ours, e_ours, vir = atoms.f.copy(), ours.eatom[:n].copy(), ours.virial.copy()
            atoms.f[:] = 0
            ref.compute(atoms, lst)
            np.testing.assert_allclose(e_ours, ref.eatom[:n], rtol=E_RTOL, atol=E_ATOL)
            assert np.abs(f_ours - atoms.f).max() < F_ATOL
            assert np.abs(vir - ref.virial).max() < V_RTOL * max(1.0, np.abs(ref.virial).max())
            assert abs(pe - ref.eng_vdwl) < 1e-5 * max(1.0, abs(ref.eng_vdwl))
        return f, pe

    f, pe = force(x, True)
    e0 = pe
    etot = []
    for step in range(1, nsteps + 1):
        v += 0.5 * dt * FTM2V * f / mass
        x += dt * v
        f, pe = force(x, step % 10 == 0)
        v += 0.5 * dt * FTM2V * f / mass
        etot.append(pe + 0.5 * MVV2E * float((mass * v * v).sum()))
    etot = np.array(etot)
    ke = etot[-1] - pe
    print("C1 NVE: E_tot(0) %.6f  E_tot(100) %.6f  KE(100) %.6f  max drift %.2e" % (e0, etot[-1], ke, np.abs(etot - e0).max()))
    assert ke > 0 and np.abs(x - pos).max() > 1e-4                          # the system actually moved
    assert np.abs(etot - e0).max() < 2e-3 * max(ke, 1e-3) + 1e-5            # velocity-Verlet conserves E to O(dt^2)


@pytest.mark.parametrize("world", [2, 4])
def test_rank_count_invariance(world, ensure_built, tmp_path):
    from lmpshim import harness as H
    from pair_allegro_b200.pair import PairAllegroB200
    from test_gpu_parity import F_ATOL
    rng = np.random.default_rng(9)
    box = 26.0
    n = 1400
    g = int(np.ceil(n ** (1 / 3)))
    grid = np.stack(np.meshgrid(*[np.arange(g)] * 3, indexing="ij"), -1).reshape(-1, 3)[:n]
    pos = (grid + 0.5) * (box / g) + rng.normal(0, 0.25, (n, 3))
    types = rng.integers(1, 3, n).astype(np.int32)
    cell = np.diag([box] * 3)
    names = ["A", "B"]
    cfg = modelgen.default_config(type_names=names, l_max=1, num_layers=2, r_max=5.0, avg_num_neighbors=30.0, seed=4)
    alg = str(tmp_path / "m.alg")
    modelgen.random_alg(cfg, alg)
    pair = PairAllegroB200(device=0, debug_mode=False)
    pair.coeff(["*", "*", alg] + names, 2)
    pair.init_style()

    def run(nranks):
        parts, rank_of, local_index = H.decompose(pos, types, cell, [True] * 3, nranks, 6.0)
        fs = [np.zeros((p.nlocal, 3)) for p in parts]
        es = [None] * nranks
        pe, vir = 0.0, np.zeros(6)
        ghost_f = []
        for r, p in enumerate(parts):
            lst = H.build_full_list(p, 6.0)
            p.f[:] = 0
            pair.compute(p, lst, eflag=1, vflag=1)
            fs[r] += p.f[:p.nlocal]
            ghost_f.append(p.f[p.nlocal:].copy())
            es[r] = pair.eatom[:p.nlocal].copy()
            pe += pair.eng_vdwl
            vir += pair.virial
        for r, p in enumerate(parts):                                       # reverse halo
            orank, oidx = p.owner_rank[p.nlocal:], p.owner_index[p.nlocal:]
            for s in range(nranks):
                sel = orank == s
                np.add.at(fs[s], oidx[sel], ghost_f[r][sel])
        F = np.zeros((n, 3)); E = np.zeros(n)
        for r in range(nranks):
            mine = np.nonzero(rank_of == r)[0]
            F[mine[np.argsort(local_index[mine])]] = fs[r]
            E[mine[np.argsort(local_index[mine])]] = es[r]
        return F, E, pe, vir

    F1, E1, pe1, v1 = run(1)
    Fn, En, pen, vn = run(world)
    assert np.abs(Fn - F1).max() < F_ATOL
    np.testing.assert_allclose(En, E1, rtol=1e-5, atol=1e-5)
    assert abs(pen - pe1) < 1e-5 * max(1.0, abs(pe1))
    assert np.abs(vn - v1).max() < 1e-4 * max(1.0, np.abs(v1).max())
    assert np.abs(F1.sum(0)).max() < 1e-4
