#!/usr/bin/env python
"""Generate the irreps / tensor-product path / real Clebsch-Gordan tables of the Allegro
network spec frozen in DESIGN.md (Variant B).

Outputs (both committed):
  tables/allegro_tables.json          -- data consumed by the oracle, the exporter and the tests
  pair_allegro_b200/csrc/tp_gen.cuh   -- fully unrolled CUDA tensor-product code (forward + backward)

The real CG tensors are computed *numerically* as the rotation-invariant tensor of
D^{l1} x D^{l2} x D^{l3} where D^l is the representation carried by the real spherical
harmonics defined below, so they are consistent with these harmonics by construction
(no dependence on any external convention).  Normalisation: ||C||_F^2 = 2*l3+1, first
non-zero entry (row-major m1,m2,m3) positive.

Self-contained on purpose: nothing under oracle/ or pair_allegro_b200/ is imported.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LMAX_SUPPORTED = 3


# ----------------------------------------------------------------------------------------
# real spherical harmonics, "component" normalisation (sum_m Y_lm^2 = 2l+1 on the sphere)
# order within l: m = -l..l  (l=1: y, z, x)
# ----------------------------------------------------------------------------------------
def real_sh(n, lmax):
    """n: [...,3] unit vectors -> [..., (lmax+1)^2]"""
    x, y, z = n[..., 0], n[..., 1], n[..., 2]
    out = [np.ones_like(x)]
    if lmax >= 1:
        s3 = np.sqrt(3.0)
        out += [s3 * y, s3 * z, s3 * x]
    if lmax >= 2:
        s15 = np.sqrt(15.0)
        s5 = np.sqrt(5.0)
        out += [s15 * x * y, s15 * y * z, 0.5 * s5 * (3 * z * z - 1.0), s15 * x * z,
                0.5 * s15 * (x * x - y * y)]
    if lmax >= 3:
        a = np.sqrt(35.0 / 8.0)
        b = np.sqrt(105.0)
        c = np.sqrt(21.0 / 8.0)
        d = 0.5 * np.sqrt(7.0)
        out += [a * y * (3 * x * x - y * y), b * x * y * z, c * y * (5 * z * z - 1.0),
                d * (5 * z * z * z - 3 * z), c * x * (5 * z * z - 1.0),
                0.5 * b * (x * x - y * y) * z, a * x * (x * x - 3 * y * y)]
    return np.stack(out, axis=-1)


def rand_rot(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def wigner_real(l, R, rng):
    """D with Y_l(R n) = D Y_l(n), from a least-squares fit on random points."""
    pts = rng.normal(size=(64, 3))
    pts /= np.linalg.norm(pts, axis=1, keepdims=True)
    sl = slice(l * l, (l + 1) * (l + 1))
    Y = real_sh(pts, l)[:, sl]            # [P, 2l+1]
    YR = real_sh(pts @ R.T, l)[:, sl]     # Y(R n)
    # YR^T = D Y^T  ->  D = (lstsq(Y, YR))^T
    D = np.linalg.lstsq(Y, YR, rcond=None)[0].T
    return D


def real_cg(l1, l2, l3, rng):
    d1, d2, d3 = 2 * l1 + 1, 2 * l2 + 1, 2 * l3 + 1
    rows = []
    for _ in range(4):
        R = rand_rot(rng)
        D1, D2, D3 = (wigner_real(l, R, rng) for l in (l1, l2, l3))
        K = np.einsum('ai,bj,ck->abcijk', D1, D2, D3).reshape(d1 * d2 * d3, d1 * d2 * d3)
        rows.append(K - np.eye(d1 * d2 * d3))
    A = np.concatenate(rows, axis=0)
    _, s, vt = np.linalg.svd(A)
    null = vt[s.size - 1]
    assert s[-1] < 1e-10, (l1, l2, l3, s[-3:])
    assert s.size == 1 or s[-2] > 1e-3, (l1, l2, l3, s[-3:])
    C = null.reshape(d1, d2, d3)
    C[np.abs(C) < 1e-10] = 0.0
    C *= np.sqrt(d3) / np.linalg.norm(C)
    first = C.reshape(-1)[np.nonzero(C.reshape(-1))[0][0]]
    if first < 0:
        C = -C
    return C


# ----------------------------------------------------------------------------------------
# irreps bookkeeping.  An irrep is (l, p) with p = +1 (even) / -1 (odd).
# ----------------------------------------------------------------------------------------
def sh_irreps(L):
    return [(l, (-1) ** l) for l in range(L + 1)]


def irrep_sort_key(ir):
    l, p = ir
    # SH-parity irreps first in l order, then the others in l order
    return (0 if p == (-1) ** l else 1, l)


def reachable(in_irreps, L):
    out = set()
    for (l1, p1) in in_irreps:
        for l2 in range(L + 1):
            p2 = (-1) ** l2
            for l3 in range(abs(l1 - l2), min(L, l1 + l2) + 1):
                out.add((l3, p1 * p2))
    return sorted(out, key=irrep_sort_key)


def useful_for(next_out_irreps, cand, L):
    """keep candidates that can reach one of next_out_irreps in one tensor product with SH."""
    keep = []
    for (l1, p1) in cand:
        ok = False
        for l2 in range(L + 1):
            p2 = (-1) ** l2
            for l3 in range(abs(l1 - l2), min(L, l1 + l2) + 1):
                if (l3, p1 * p2) in next_out_irreps:
                    ok = True
        if ok:
            keep.append((l1, p1))
    return keep


def offsets(irreps):
    off, o = [], 0
    for (l, _) in irreps:
        off.append(o)
        o += 2 * l + 1
    return off, o


def build_kind(in_irreps, out_irreps, L, cg):
    """paths (l1,p1) x (l2,(-1)^l2) -> (l3,p3) in out_irreps; sorted by (out idx, in idx, l2)."""
    in_off, din = offsets(in_irreps)
    out_off, dout = offsets(out_irreps)
    paths = []
    for o3, (l3, p3) in enumerate(out_irreps):
        for i1, (l1, p1) in enumerate(in_irreps):
            for l2 in range(L + 1):
                p2 = (-1) ** l2
                if p1 * p2 != p3 or not (abs(l1 - l2) <= l3 <= l1 + l2):
                    continue
                C = cg[(l1, l2, l3)]
                nz = [[int(a), int(b), int(c), float(C[a, b, c])]
                      for a in range(2 * l1 + 1) for b in range(2 * l2 + 1) for c in range(2 * l3 + 1)
                      if C[a, b, c] != 0.0]
                paths.append(dict(l1=l1, p1=p1, i1=i1, in_off=in_off[i1], l2=l2, sh_off=l2 * l2,
                                  l3=l3, p3=p3, o3=o3, out_off=out_off[o3],
                                  scalar=bool(l3 == 0 and p3 == 1), nz=nz))
    n0 = sum(1 for p in paths if p['scalar'])
    return dict(in_irreps=[list(i) for i in in_irreps], out_irreps=[list(i) for i in out_irreps],
                din=din, dout=dout, n_paths=len(paths), n0=n0, paths=paths)


def build_tables():
    rng = np.random.default_rng(12345)
    cg = {}
    for l1 in range(LMAX_SUPPORTED + 1):
        for l2 in range(LMAX_SUPPORTED + 1):
            for l3 in range(abs(l1 - l2), min(LMAX_SUPPORTED, l1 + l2) + 1):
                cg[(l1, l2, l3)] = real_cg(l1, l2, l3, rng)
    tables = {"format": 1, "lmax_supported": LMAX_SUPPORTED, "L": {}}
    for L in range(1, LMAX_SUPPORTED + 1):
        SH = sh_irreps(L)
        scal = [(0, 1)]
        # FULL1: reachable from SH x SH and useful for a following SH-output layer
        FULL1 = useful_for(SH, reachable(SH, L), L)
        kinds = {
            # last layer of any net: only scalar outputs (s_e); no V_out
            "A": build_kind(SH, scal, L, cg),
            # first layer of a 2-layer net: SH -> SH
            "B": build_kind(SH, SH, L, cg),
            # first layer of a 3-layer net: SH -> FULL1
            "C": build_kind(SH, FULL1, L, cg),
            # middle layer of a 3-layer net: FULL1 -> SH
            "D": build_kind(FULL1, SH, L, cg),
        }
        tables["L"][str(L)] = dict(nsh=(L + 1) ** 2, sh_irreps=[list(i) for i in SH],
                                   full1_irreps=[list(i) for i in FULL1], kinds=kinds)
    return tables


def layer_kinds(n_layers):
    return {1: ["A"], 2: ["B", "A"], 3: ["C", "D", "A"]}[n_layers]


# ----------------------------------------------------------------------------------------
# CUDA code generation
# ----------------------------------------------------------------------------------------
def fl(c):
    return repr(np.float32(c).item()) + "f" if "." in repr(np.float32(c).item()) or "e" in repr(np.float32(c).item()) else repr(np.float32(c).item()) + ".0f"


def gen_cuda(tables):
    o = []
    w = o.append
    w("// GENERATED by tools/gen_tables.py -- do not edit.  Fully unrolled channel-wise")
    w("// Clebsch-Gordan tensor products (one channel u of one edge per call).")
    w("//   tp_fwd : T[path] = sum C * Vin[l1] * G[l2];  s[k] = scalar-path outputs (unmixed);")
    w("//            Vout[(l3,p3)] = sum_paths omega[path] * T[path]")
    w("//   tp_bwd : given dVout, ds -> dVin, dG   (omega, Vin, G as in forward)")
    w("// omega is indexed omega[path * OMEGA_STRIDE] (stride = number of channels U).")
    w("#pragma once")
    w("namespace tpgen {")
    w("template <int L, char KIND> struct TP;")
    for Ls, TL in tables["L"].items():
        for kname, K in TL["kinds"].items():
            w(f"template <> struct TP<{Ls}, '{kname}'> {{")
            w(f"  static constexpr int DIN = {K['din']}, DOUT = {K['dout']}, NPATH = {K['n_paths']}, N0 = {K['n0']}, NSH = {TL['nsh']};")
            has_vout = kname != "A"
            # ---------------- forward
            w("  template <int OS> __device__ __forceinline__ static void fwd(const float* __restrict__ Vin, const float* __restrict__ G,")
            w("      const float* __restrict__ omega, float* __restrict__ Vout, float* __restrict__ s) {")
            if has_vout:
                w(f"    #pragma unroll\n    for (int i = 0; i < {K['dout']}; ++i) Vout[i] = 0.f;")
            sidx = 0
            for pi, P in enumerate(K["paths"]):
                d3 = 2 * P["l3"] + 1
                w(f"    {{ // path {pi}: ({P['l1']},{P['p1']:+d}) x {P['l2']} -> ({P['l3']},{P['p3']:+d})")
                terms = [[] for _ in range(d3)]
                for (a, b, c, v) in P["nz"]:
                    terms[c].append(f"{fl(v)} * Vin[{P['in_off'] + a}] * G[{P['sh_off'] + b}]")
                for c in range(d3):
                    w(f"      const float t{c} = " + (" + ".join(terms[c]) if terms[c] else "0.f") + ";")
                if P["scalar"]:
                    w(f"      s[{sidx}] = t0;")
                    sidx += 1
                if has_vout:
                    w(f"      const float om = omega[{pi} * OS];")
                    for c in range(d3):
                        w(f"      Vout[{P['out_off'] + c}] += om * t{c};")
                w("    }")
            w("  }")
            # ---------------- backward
            w("  template <int OS> __device__ __forceinline__ static void bwd(const float* __restrict__ Vin, const float* __restrict__ G,")
            w("      const float* __restrict__ omega, const float* __restrict__ dVout, const float* __restrict__ ds,")
            w("      float* __restrict__ dVin, float* __restrict__ dG) {")
            w(f"    #pragma unroll\n    for (int i = 0; i < {K['din']}; ++i) dVin[i] = 0.f;")
            w(f"    #pragma unroll\n    for (int i = 0; i < {TL['nsh']}; ++i) dG[i] = 0.f;")
            sidx = 0
            for pi, P in enumerate(K["paths"]):
                d3 = 2 * P["l3"] + 1
                w(f"    {{ // path {pi}")
                if has_vout:
                    w(f"      const float om = omega[{pi} * OS];")
                for c in range(d3):
                    expr = f"om * dVout[{P['out_off'] + c}]" if has_vout else None
                    if P["scalar"]:
                        expr = (expr + f" + ds[{sidx}]") if expr else f"ds[{sidx}]"
                    w(f"      const float d{c} = {expr};")
                if P["scalar"]:
                    sidx += 1
                for (a, b, c, v) in P["nz"]:
                    w(f"      dVin[{P['in_off'] + a}] += {fl(v)} * d{c} * G[{P['sh_off'] + b}];")
                    w(f"      dG[{P['sh_off'] + b}] += {fl(v)} * d{c} * Vin[{P['in_off'] + a}];")
                w("    }")
            w("  }")
            w("};")
    w("}  // namespace tpgen")
    return "\n".join(o) + "\n"


def main():
    tables = build_tables()
    os.makedirs(os.path.join(ROOT, "tables"), exist_ok=True)
    jpath = os.path.join(ROOT, "tables", "allegro_tables.json")
    with open(jpath, "w") as f:
        json.dump(tables, f, separators=(",", ":"))
    cpath = os.path.join(ROOT, "pair_allegro_b200", "csrc", "tp_gen.cuh")
    with open(cpath, "w") as f:
        f.write(gen_cuda(tables))
    for Ls, TL in tables["L"].items():
        for k, K in TL["kinds"].items():
            nnz = sum(len(p["nz"]) for p in K["paths"])
            print(f"L={Ls} kind {k}: in {K['in_irreps']} out {K['out_irreps']} din={K['din']} dout={K['dout']} "
                  f"paths={K['n_paths']} n0={K['n0']} nnz={nnz}")
    print("wrote", jpath, cpath)


if __name__ == "__main__":
    sys.exit(main())
