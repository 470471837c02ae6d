"""Stand-in for the parts of LAMMPS core the pair style reads: the CALLER side of the boundary
(test / bench infrastructure; LAMMPS itself is not in /root/reference and is out of scope).
No model arithmetic lives here.

Provides, in numpy, exactly the state `PairNequIPAllegro::preprocess()` consumes
(/root/reference/pair_nequip_allegro.cpp:459-480): atom->x/type/tag/nlocal with periodic
ghost images, and a FULL neighbour list (inum, ilist, numneigh, firstneigh) of the local
atoms built with cutoff r_max + skin (`neighbor 1.0 bin`,
/root/reference/tests/test_python_repro_allegro.py:100).  Also: extended-xyz reader for the
reference's geometry fixtures, brick domain decomposition for N "ranks", synthetic boxes of
BASELINE.json's configs.
"""
import re
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

NEIGHMASK = 0x1FFFFFFF  # LAMMPS lmptype.h: low 29 bits hold the atom index


@dataclass
class Atoms:
    """mirror of the LAMMPS `Atom` fields the pair style touches"""
    x: np.ndarray            # [nlocal+nghost,3] f64
    type: np.ndarray         # [ntot] int32, 1-based LAMMPS types
    tag: np.ndarray          # [ntot] int64, 1-based atom IDs (ghosts carry the owner's tag)
    nlocal: int
    nghost: int
    ntypes: int
    f: np.ndarray = None     # [ntot,3] f64
    owner: np.ndarray = None  # [ntot] index of the owning local atom (single rank) for reverse comm
    # multi-rank only:
    owner_rank: np.ndarray = None  # [ntot] rank owning each atom (ghosts)
    owner_index: np.ndarray = None  # [ntot] local index on the owning rank

    def __post_init__(self):
        if self.f is None:
            self.f = np.zeros_like(self.x)


@dataclass
class NeighList:
    """mirror of LAMMPS `NeighList` (full list of local atoms)"""
    inum: int
    gnum: int
    ilist: np.ndarray        # [inum+gnum] int32
    numneigh: np.ndarray     # [ntot] int32 (0 for ghosts)
    neigh_flat: np.ndarray   # concatenated neighbour indices (int32), atom i's slice = first[i]:first[i]+numneigh[i]
    first: np.ndarray        # [ntot] int64 offsets into neigh_flat

    def firstneigh(self, i):
        return self.neigh_flat[self.first[i]:self.first[i] + self.numneigh[i]]


def read_extxyz_frame(path, frame=0):
    """minimal extended-xyz reader (species + pos; Lattice, pbc from the comment line)."""
    with open(path) as f:
        lines = f.read().splitlines()
    p = 0
    for _ in range(frame + 1):
        n = int(lines[p].strip())
        comment = lines[p + 1]
        body = lines[p + 2:p + 2 + n]
        p += 2 + n
    m = re.search(r'Lattice="([^"]+)"', comment)
    cell = np.array([float(v) for v in m.group(1).split()]).reshape(3, 3) if m else None
    m = re.search(r'pbc="([^"]+)"', comment)
    pbc = [t.upper().startswith("T") for t in m.group(1).split()] if m else [False] * 3
    species = [b.split()[0] for b in body]
    pos = np.array([[float(v) for v in b.split()[1:4]] for b in body])
    return species, pos, cell, pbc


def _image_shifts(cell, pbc, rcomm):
    """integer image ranges so that every image within rcomm of the cell is covered."""
    inv = np.linalg.inv(cell)
    # distance between the two faces perpendicular to reciprocal vector k = 1/|inv[:,k]|
    heights = 1.0 / np.linalg.norm(inv, axis=0)
    return [int(np.ceil(rcomm / h)) if pb else 0 for h, pb in zip(heights, pbc)]


def make_single_rank(species_or_types, pos, cell, pbc, rcomm, type_names=None):
    """one-rank LAMMPS picture: wrapped local atoms + all periodic images within rcomm of the box
    (lamda coordinates in [-rcomm/h, 1+rcomm/h)), in the order LAMMPS would append them
    (images after locals)."""
    pos = np.asarray(pos, dtype=np.float64)
    n = len(pos)
    if type_names is not None:
        types = np.array([type_names.index(s) + 1 for s in species_or_types], dtype=np.int32)
        ntypes = len(type_names)
    else:
        types = np.asarray(species_or_types, dtype=np.int32)
        ntypes = int(types.max())
    cell = np.asarray(cell, dtype=np.float64)
    inv = np.linalg.inv(cell)
    lam = pos @ inv
    for k in range(3):
        if pbc[k]:
            lam[:, k] -= np.floor(lam[:, k])
    pos = lam @ cell
    heights = 1.0 / np.linalg.norm(inv, axis=0)
    nimg = _image_shifts(cell, pbc, rcomm)
    xs, ts, tg, own = [pos], [types], [np.arange(1, n + 1, dtype=np.int64)], [np.arange(n)]
    margin = rcomm / heights
    for a in range(-nimg[0], nimg[0] + 1):
        for b in range(-nimg[1], nimg[1] + 1):
            for c in range(-nimg[2], nimg[2] + 1):
                if a == 0 and b == 0 and c == 0:
                    continue
                sh = np.array([a, b, c], dtype=np.float64)
                l2 = lam + sh
                keep = np.all((l2 >= -margin) & (l2 < 1.0 + margin), axis=1)
                if not keep.any():
                    continue
                xs.append(pos[keep] + sh @ cell)
                ts.append(types[keep])
                tg.append(np.arange(1, n + 1, dtype=np.int64)[keep])
                own.append(np.arange(n)[keep])
    x = np.concatenate(xs)
    return Atoms(x=x, type=np.concatenate(ts), tag=np.concatenate(tg), nlocal=n, nghost=len(x) - n,
                 ntypes=ntypes, owner=np.concatenate(own))


def build_full_list(atoms: Atoms, rneigh: float) -> NeighList:
    """FULL neighbour list of the local atoms over locals+ghosts, |dx|^2 <= rneigh^2, j != i,
    ascending j within each row (any order is a legal LAMMPS list).  Cell-list, vectorised."""
    x = atoms.x
    ntot = len(x)
    nl = atoms.nlocal
    lo = x.min(axis=0) - 1e-9
    binsz = max(rneigh, 1e-6)
    nb = np.maximum(((x.max(axis=0) - lo) / binsz).astype(np.int64) + 1, 1)
    bi = np.minimum(((x - lo) / binsz).astype(np.int64), nb - 1)
    key = (bi[:, 0] * nb[1] + bi[:, 1]) * nb[2] + bi[:, 2]
    order = np.argsort(key, kind="stable")
    skey = key[order]
    nbins = int(nb.prod())
    start = np.searchsorted(skey, np.arange(nbins + 1))
    pairs_i, pairs_j = [], []
    loc = np.arange(nl)
    r2 = rneigh * rneigh
    for da in (-1, 0, 1):
        for db in (-1, 0, 1):
            for dc in (-1, 0, 1):
                b2 = bi[:nl] + np.array([da, db, dc])
                ok = np.all((b2 >= 0) & (b2 < nb), axis=1)
                k2 = (b2[:, 0] * nb[1] + b2[:, 1]) * nb[2] + b2[:, 2]
                k2 = np.where(ok, k2, 0)
                s, e = start[k2], start[k2 + 1]
                cnt = np.where(ok, e - s, 0)
                if cnt.sum() == 0:
                    continue
                ii = np.repeat(loc, cnt)
                off = np.arange(cnt.sum()) - np.repeat(np.cumsum(cnt) - cnt, cnt)
                jj = order[np.repeat(s, cnt) + off]
                d = x[ii] - x[jj]
                m = ((d * d).sum(axis=1) <= r2) & (ii != jj)
                pairs_i.append(ii[m])
                pairs_j.append(jj[m])
    ii = np.concatenate(pairs_i) if pairs_i else np.zeros(0, dtype=np.int64)
    jj = np.concatenate(pairs_j) if pairs_j else np.zeros(0, dtype=np.int64)
    o = np.lexsort((jj, ii))
    ii, jj = ii[o], jj[o]
    numneigh = np.zeros(ntot, dtype=np.int32)
    numneigh[:nl] = np.bincount(ii, minlength=nl)[:nl]
    first = np.zeros(ntot, dtype=np.int64)
    first[1:] = np.cumsum(numneigh)[:-1]
    return NeighList(inum=nl, gnum=atoms.nghost, ilist=np.arange(ntot, dtype=np.int32), numneigh=numneigh,
                     neigh_flat=jj.astype(np.int32), first=first)


def reverse_comm_single_rank(atoms: Atoms, f: np.ndarray) -> np.ndarray:
    """what LAMMPS `comm->reverse_comm()` does after the pair style with newton on: ghost
    forces are added to their owners (here all owners are local)."""
    out = f[:atoms.nlocal].copy()
    np.add.at(out, atoms.owner[atoms.nlocal:], f[atoms.nlocal:])
    return out


# ----------------------------------------------------------------------------------------
# synthetic boxes (BASELINE.json configs)
# ----------------------------------------------------------------------------------------
def fcc_box(ncell, a=4.09, jitter=0.05, seed=2):
    """config C2: FCC, `ncell`^3 conventional cells, N(0,jitter) displacement per coordinate."""
    rng = np.random.default_rng(seed)
    base = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5]])
    g = np.stack(np.meshgrid(np.arange(ncell), np.arange(ncell), np.arange(ncell), indexing="ij"), -1).reshape(-1, 3)
    pos = (g[:, None, :] + base[None]).reshape(-1, 3) * a
    pos = pos + rng.normal(0.0, jitter, pos.shape)
    cell = np.eye(3) * ncell * a
    return pos, np.ones(len(pos), dtype=np.int32), cell


def water_like_box(nmol, density=0.1, seed=3):
    """config C3: O on a jittered simple-cubic lattice, two H per O at 0.96 A / 104.5 deg,
    random orientation.  types: 1 = H, 2 = O."""
    rng = np.random.default_rng(seed)
    n = nmol * 3
    box = (n / density) ** (1.0 / 3.0)
    m = int(np.ceil(nmol ** (1.0 / 3.0)))
    g = np.stack(np.meshgrid(np.arange(m), np.arange(m), np.arange(m), indexing="ij"), -1).reshape(-1, 3)[:nmol]
    O = (g + 0.5) * (box / m) + rng.normal(0.0, 0.3, (nmol, 3))
    q = rng.normal(size=(nmol, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q.T
    R = np.stack([np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], -1),
                  np.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], -1),
                  np.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1)], 1)
    th = np.deg2rad(104.5) / 2
    h1 = 0.96 * np.array([np.sin(th), 0, np.cos(th)])
    h2 = 0.96 * np.array([-np.sin(th), 0, np.cos(th)])
    H1 = O + R @ h1
    H2 = O + R @ h2
    pos = np.concatenate([O, H1, H2])
    types = np.concatenate([np.full(nmol, 2), np.full(2 * nmol, 1)]).astype(np.int32)
    return pos, types, np.eye(3) * box


def multi_species_box(natoms, fractions=(3, 1, 4, 0), density=0.09, seed=5, jitter=0.15):
    """config C5: Li3PO4-like jittered lattice, up to 4 species (counts ~ fractions)."""
    rng = np.random.default_rng(seed)
    box = (natoms / density) ** (1.0 / 3.0)
    m = int(np.ceil(natoms ** (1.0 / 3.0)))
    g = np.stack(np.meshgrid(np.arange(m), np.arange(m), np.arange(m), indexing="ij"), -1).reshape(-1, 3)
    sel = rng.permutation(len(g))[:natoms]
    pos = (g[sel] + 0.5) * (box / m) + rng.normal(0.0, jitter, (natoms, 3))
    fr = np.array(fractions, dtype=np.float64)
    fr = fr / fr.sum()
    types = (rng.choice(len(fr), size=natoms, p=fr) + 1).astype(np.int32)
    return pos, types, np.eye(3) * box


# ----------------------------------------------------------------------------------------
# brick domain decomposition (what LAMMPS Comm does for N ranks), orthogonal boxes
# ----------------------------------------------------------------------------------------
def proc_grid(nranks):
    return {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[nranks]


def decompose(pos, types, cell, pbc, nranks, rcomm):
    """Split a periodic orthogonal box into a brick grid.  Returns a list of per-rank `Atoms`
    whose ghosts carry (owner_rank, owner_index) so forward/reverse halo exchange can be
    executed by any transport (numpy in CPU tests, NCCL in bench.py)."""
    L = np.diag(cell).astype(np.float64)
    assert np.allclose(cell, np.diag(L)), "brick decomposition: orthogonal boxes only"
    pos = np.asarray(pos, dtype=np.float64).copy()
    for k in range(3):
        if pbc[k]:
            pos[:, k] -= np.floor(pos[:, k] / L[k]) * L[k]
    grid = np.array(proc_grid(nranks))
    sub = L / grid
    cidx = np.minimum((pos / sub).astype(np.int64), grid - 1)
    rank_of = (cidx[:, 0] * grid[1] + cidx[:, 1]) * grid[2] + cidx[:, 2]
    order = np.argsort(rank_of, kind="stable")
    counts = np.bincount(rank_of, minlength=nranks)
    starts = np.concatenate([[0], np.cumsum(counts)])
    local_index = np.empty(len(pos), dtype=np.int64)
    local_index[order] = np.arange(len(pos)) - starts[rank_of[order]]
    gtag = np.arange(1, len(pos) + 1, dtype=np.int64)
    ntypes = int(types.max())
    out = []
    nimg = [int(np.ceil(rcomm / L[k])) if pbc[k] else 0 for k in range(3)]
    for r in range(nranks):
        c = np.array([r // (grid[1] * grid[2]), (r // grid[2]) % grid[1], r % grid[2]])
        lo, hi = c * sub, (c + 1) * sub
        mine = order[starts[r]:starts[r + 1]]
        xs, ts, tg, orank, oidx = [pos[mine]], [types[mine]], [gtag[mine]], [np.full(len(mine), r)], [local_index[mine]]
        for a in range(-nimg[0], nimg[0] + 1):
            for b in range(-nimg[1], nimg[1] + 1):
                for cc in range(-nimg[2], nimg[2] + 1):
                    sh = np.array([a, b, cc]) * L
                    p2 = pos + sh
                    keep = np.all((p2 >= lo - rcomm) & (p2 < hi + rcomm), axis=1)
                    if a == 0 and b == 0 and cc == 0:
                        keep &= rank_of != r
                    if not keep.any():
                        continue
                    xs.append(p2[keep]); ts.append(types[keep]); tg.append(gtag[keep])
                    orank.append(rank_of[keep]); oidx.append(local_index[keep])
        x = np.concatenate(xs)
        out.append(Atoms(x=x, type=np.concatenate(ts).astype(np.int32), tag=np.concatenate(tg), nlocal=len(mine),
                         nghost=len(x) - len(mine), ntypes=ntypes,
                         owner_rank=np.concatenate(orank), owner_index=np.concatenate(oidx)))
    return out, rank_of, local_index


def decompose_rank(pos, types, cell, pbc, nranks, rank, rcomm):
    """One rank's view of the brick decomposition plus its halo plan.

    Ghosts are ordered by (owner rank, image shift id, owner local index) so that the ghosts
    owned by one peer form a contiguous slice (forward comm can receive straight into x, reverse
    comm can send straight out of f) and so that the owner can reproduce the order on its own.
    Returns (atoms, plan) with plan = dict(
        recv_slices[s] = (start, stop) ghost index range owned by rank s (absent if empty),
        send_index[s]  = local atom indices this rank ships to rank s (in s's ghost order),
        send_shift[s]  = [n,3] shift to add to x when packing for rank s)."""
    L = np.diag(cell).astype(np.float64)
    assert np.allclose(cell, np.diag(L)), "brick decomposition: orthogonal boxes only"
    pos = np.asarray(pos, dtype=np.float64).copy()
    for k in range(3):
        if pbc[k]:
            pos[:, k] -= np.floor(pos[:, k] / L[k]) * L[k]
    grid = np.array(proc_grid(nranks))
    sub = L / grid
    cidx = np.minimum((pos / sub).astype(np.int64), grid - 1)
    rank_of = (cidx[:, 0] * grid[1] + cidx[:, 1]) * grid[2] + cidx[:, 2]
    mine = np.nonzero(rank_of == rank)[0]
    counts = np.bincount(rank_of, minlength=nranks)
    # local index of every atom on its owner = rank within the owner's (stable, global-order) list
    order = np.argsort(rank_of, kind="stable")
    starts = np.concatenate([[0], np.cumsum(counts)])
    local_index = np.empty(len(pos), dtype=np.int64)
    local_index[order] = np.arange(len(pos)) - starts[rank_of[order]]
    shifts = [(a, b, c) for a in ((-1, 0, 1) if pbc[0] else (0,)) for b in ((-1, 0, 1) if pbc[1] else (0,))
              for c in ((-1, 0, 1) if pbc[2] else (0,))]
    assert np.all(sub >= rcomm), "sub-domain thinner than the ghost cutoff"

    def brick(r):
        c = np.array([r // (grid[1] * grid[2]), (r // grid[2]) % grid[1], r % grid[2]])
        return c * sub, (c + 1) * sub

    lo, hi = brick(rank)
    # --- my ghosts
    g_idx, g_shift, g_sid = [], [], []
    for sid, sh in enumerate(shifts):
        shv = np.array(sh) * L
        p2 = pos + shv
        keep = np.all((p2 >= lo - rcomm) & (p2 < hi + rcomm), axis=1)
        if sh == (0, 0, 0):
            keep &= rank_of != rank
        idx = np.nonzero(keep)[0]
        g_idx.append(idx); g_shift.append(np.tile(shv, (len(idx), 1))); g_sid.append(np.full(len(idx), sid))
    g_idx = np.concatenate(g_idx); g_shift = np.concatenate(g_shift); g_sid = np.concatenate(g_sid)
    o = np.lexsort((local_index[g_idx], g_sid, rank_of[g_idx]))
    g_idx, g_shift = g_idx[o], g_shift[o]
    x = np.concatenate([pos[mine], pos[g_idx] + g_shift])
    atoms = Atoms(x=x, type=np.concatenate([types[mine], types[g_idx]]).astype(np.int32),
                  tag=np.concatenate([mine, g_idx]).astype(np.int64) + 1, nlocal=len(mine), nghost=len(g_idx),
                  ntypes=int(types.max()), owner_rank=np.concatenate([np.full(len(mine), rank), rank_of[g_idx]]),
                  owner_index=np.concatenate([np.arange(len(mine)), local_index[g_idx]]))
    plan = dict(recv_slices={}, send_index={}, send_shift={})
    gowner = rank_of[g_idx]
    for s in range(nranks):
        w = np.nonzero(gowner == s)[0]
        if len(w):
            plan["recv_slices"][s] = (len(mine) + int(w[0]), len(mine) + int(w[-1]) + 1)
    # --- what I ship to each peer s: my atoms inside s's extended brick, (shift id, local index) order
    mypos = pos[mine]
    for s in range(nranks):
        slo, shi = brick(s)
        si, ss = [], []
        for sh in shifts:
            if s == rank and sh == (0, 0, 0):
                continue
            shv = np.array(sh) * L
            p2 = mypos + shv
            keep = np.all((p2 >= slo - rcomm) & (p2 < shi + rcomm), axis=1)
            idx = np.nonzero(keep)[0]
            si.append(idx); ss.append(np.tile(shv, (len(idx), 1)))
        si = np.concatenate(si); ss = np.concatenate(ss)
        if len(si):
            plan["send_index"][s] = si.astype(np.int32)
            plan["send_shift"][s] = ss
    return atoms, plan
