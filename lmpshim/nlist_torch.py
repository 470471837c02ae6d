"""Binned FULL neighbour-list build with torch tensor ops (CUDA when available, CPU otherwise) -- the step LAMMPS'
`Neighbor` class performs before the pair style ("neighbour-list construction on GPU").
Caller-side infrastructure like the rest of lmpshim/: bench.py and the full-size tests use it so that setting up a
1 M ... 8 M-atom box costs seconds instead of minutes of numpy; no model arithmetic lives here.

Same contract as `harness.build_full_list`: for every LOCAL atom i all atoms j != i (locals + ghosts) with
|x_i - x_j|^2 <= rneigh^2.  The order inside a row is (cell offset, position in the cell) -- any order is a legal LAMMPS
list; `tests/test_nlist_torch.py` checks the rows against the numpy builder as sets."""
import numpy as np
import torch


def build_full_list_torch(x, nlocal, rneigh, device=None, chunk=32768, want_host=True, want_2d=True):
    """x: [ntot,3] float64 (numpy).  Returns a dict with
         numneigh [nlocal] int32 (torch, on `device`),
         nb2d [nlocal, maxn] int32 (torch, on `device`; the Kokkos-style d_neighbors(i,jj) view, LayoutRight) if want_2d,
         numneigh_h / neigh_flat_h / first_h (numpy CSR, the host list alg_compute_host takes) if want_host."""
    dev = torch.device(device if device is not None else ("cuda" if torch.cuda.is_available() else "cpu"))
    xt = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).to(dev)
    ntot = xt.shape[0]
    lo = xt.min(0).values - 1e-9
    binsz = max(float(rneigh), 1e-6)
    nb = torch.clamp(((xt.max(0).values - lo) / binsz).long() + 1, min=1)
    bi = torch.minimum(((xt - lo) / binsz).long(), nb - 1)
    key = (bi[:, 0] * nb[1] + bi[:, 1]) * nb[2] + bi[:, 2]
    order = torch.argsort(key, stable=True)
    skey = key[order]
    nbins = int(nb.prod().item())
    start = torch.searchsorted(skey, torch.arange(nbins + 1, device=dev))
    m = int((start[1:] - start[:-1]).max().item())
    off = torch.tensor([[a, b, c] for a in (-1, 0, 1) for b in (-1, 0, 1) for c in (-1, 0, 1)], device=dev)
    ar = torch.arange(m, device=dev)
    r2 = float(rneigh) * float(rneigh)
    rows_cnt, rows_j = [], []
    chunk = max(1024, int(chunk * 30 / max(m, 1)))            # keep the [n, 27, m] candidate block bounded
    for c0 in range(0, nlocal, chunk):
        c1 = min(nlocal, c0 + chunk)
        b = bi[c0:c1]
        nbins3 = b[:, None, :] + off[None]
        ok = ((nbins3 >= 0) & (nbins3 < nb)).all(-1)
        k2 = torch.clamp((nbins3[..., 0] * nb[1] + nbins3[..., 1]) * nb[2] + nbins3[..., 2], 0, nbins - 1)
        s = start[k2]
        e = torch.where(ok, start[k2 + 1], s)
        slot = s[..., None] + ar
        valid = slot < e[..., None]
        j = order[torch.clamp(slot, max=ntot - 1)]
        d = xt[c0:c1, None, None, :] - xt[j]
        keep = valid & ((d * d).sum(-1) <= r2) & (j != torch.arange(c0, c1, device=dev)[:, None, None])
        rows_cnt.append(keep.sum((1, 2)).to(torch.int32))
        rows_j.append(j[keep].to(torch.int32))                # row-major: grouped by atom, (cell offset, slot) inside
        del nbins3, ok, k2, s, e, slot, valid, j, d, keep
    numneigh = torch.cat(rows_cnt) if rows_cnt else torch.zeros(0, dtype=torch.int32, device=dev)
    flat = torch.cat(rows_j) if rows_j else torch.zeros(0, dtype=torch.int32, device=dev)
    first = torch.cumsum(numneigh.long(), 0) - numneigh.long()
    out = {"numneigh": numneigh, "candidates": int(flat.numel())}
    if want_2d:
        maxn = int(numneigh.max().item()) if nlocal else 0
        nb2d = torch.zeros(nlocal, max(maxn, 1), dtype=torch.int32, device=dev)
        rows = torch.repeat_interleave(torch.arange(nlocal, device=dev), numneigh.long())
        cols = torch.arange(flat.numel(), device=dev) - first[rows]
        nb2d[rows, cols] = flat
        out["nb2d"], out["maxn"] = nb2d, max(maxn, 1)
        del rows, cols
    if want_host:
        nn = np.zeros(ntot, dtype=np.int32)
        nn[:nlocal] = numneigh.cpu().numpy()
        fh = np.zeros(ntot, dtype=np.int64)
        fh[:nlocal] = first.cpu().numpy()
        if ntot > nlocal:
            fh[nlocal:] = int(flat.numel())
        out["numneigh_h"], out["neigh_flat_h"], out["first_h"] = nn, flat.cpu().numpy(), fh
    return out


def as_neighlist(atoms, res):
    """wrap the host CSR of build_full_list_torch as a harness.NeighList"""
    from .harness import NeighList
    ntot = atoms.nlocal + atoms.nghost
    return NeighList(inum=atoms.nlocal, gnum=atoms.nghost, ilist=np.arange(ntot, dtype=np.int32), numneigh=res["numneigh_h"],
                     neigh_flat=res["neigh_flat_h"], first=res["first_h"])
