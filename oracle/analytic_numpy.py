"""ORACLE (test infrastructure).  numpy fp64 restatement of the network of
oracle/allegro_torch.py with the HAND-DERIVED analytic backward (no autograd) in exactly
the staging the CUDA kernels use (csrc/allegro_kernels.cuh: F0 -> F_k -> T -> B_k -> B0).
tests/test_oracle.py checks it against torch autograd; it is the executable statement of
the chain rule documented in DESIGN.md and returns every intermediate the GPU tests compare
(x^k, Gamma_k, V^k, E_e, dX, dGamma, dY, du, g_e).
"""
import json
import math
import os

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ACT_C = 1.6765324703310909


def layer_kinds(n):
    return {1: ["A"], 2: ["B", "A"], 3: ["C", "D", "A"]}[n]


def sh_and_grad(n, L):
    """Y[E,nsh] and gradient wrt the (unconstrained) components of n: dY[E,nsh,3]."""
    x, y, z = n[:, 0], n[:, 1], n[:, 2]
    E = len(x)
    nsh = (L + 1) ** 2
    Y = np.zeros((E, nsh))
    G = np.zeros((E, nsh, 3))
    Y[:, 0] = 1.0
    if L >= 1:
        s3 = math.sqrt(3.0)
        Y[:, 1], Y[:, 2], Y[:, 3] = s3 * y, s3 * z, s3 * x
        G[:, 1, 1] = s3; G[:, 2, 2] = s3; G[:, 3, 0] = s3
    if L >= 2:
        s15, s5 = math.sqrt(15.0), math.sqrt(5.0)
        Y[:, 4] = s15 * x * y; G[:, 4, 0] = s15 * y; G[:, 4, 1] = s15 * x
        Y[:, 5] = s15 * y * z; G[:, 5, 1] = s15 * z; G[:, 5, 2] = s15 * y
        Y[:, 6] = 0.5 * s5 * (3 * z * z - 1); G[:, 6, 2] = 3 * s5 * z
        Y[:, 7] = s15 * x * z; G[:, 7, 0] = s15 * z; G[:, 7, 2] = s15 * x
        Y[:, 8] = 0.5 * s15 * (x * x - y * y); G[:, 8, 0] = s15 * x; G[:, 8, 1] = -s15 * y
    if L >= 3:
        a, b, c, d = math.sqrt(35 / 8), math.sqrt(105.0), math.sqrt(21 / 8), 0.5 * math.sqrt(7.0)
        Y[:, 9] = a * y * (3 * x * x - y * y); G[:, 9, 0] = 6 * a * x * y; G[:, 9, 1] = a * (3 * x * x - 3 * y * y)
        Y[:, 10] = b * x * y * z; G[:, 10, 0] = b * y * z; G[:, 10, 1] = b * x * z; G[:, 10, 2] = b * x * y
        Y[:, 11] = c * y * (5 * z * z - 1); G[:, 11, 1] = c * (5 * z * z - 1); G[:, 11, 2] = 10 * c * y * z
        Y[:, 12] = d * (5 * z ** 3 - 3 * z); G[:, 12, 2] = d * (15 * z * z - 3)
        Y[:, 13] = c * x * (5 * z * z - 1); G[:, 13, 0] = c * (5 * z * z - 1); G[:, 13, 2] = 10 * c * x * z
        Y[:, 14] = 0.5 * b * (x * x - y * y) * z; G[:, 14, 0] = b * x * z; G[:, 14, 1] = -b * y * z; G[:, 14, 2] = 0.5 * b * (x * x - y * y)
        Y[:, 15] = a * x * (x * x - 3 * y * y); G[:, 15, 0] = a * (3 * x * x - 3 * y * y); G[:, 15, 1] = -6 * a * x * y
    return Y, G


def silu(z):
    s = 1.0 / (1.0 + np.exp(-z))
    return ACT_C * z * s, ACT_C * s * (1.0 + z * (1.0 - s))


def tp_fwd(K, Vin, G, omega, want_vout):
    """Vin [E,din,U], G [E,nsh,U] -> Vout [E,dout,U] (or None), s [E,n0,U]"""
    E, _, U = Vin.shape
    Vout = np.zeros((E, K["dout"], U)) if want_vout else None
    s = []
    for ip, P in enumerate(K["paths"]):
        t = np.zeros((E, 2 * P["l3"] + 1, U))
        for (a, b, c, v) in P["nz"]:
            t[:, c] += v * Vin[:, P["in_off"] + a] * G[:, P["sh_off"] + b]
        if P["scalar"]:
            s.append(t[:, 0])
        if want_vout:
            Vout[:, P["out_off"]:P["out_off"] + 2 * P["l3"] + 1] += omega[ip][None, None, :] * t
    return Vout, np.stack(s, axis=1)


def tp_bwd(K, Vin, G, omega, dVout, ds):
    dVin = np.zeros_like(Vin)
    dG = np.zeros_like(G)
    k = 0
    for ip, P in enumerate(K["paths"]):
        d3 = 2 * P["l3"] + 1
        d = np.zeros((Vin.shape[0], d3, Vin.shape[2]))
        if dVout is not None:
            d += omega[ip][None, None, :] * dVout[:, P["out_off"]:P["out_off"] + d3]
        if P["scalar"]:
            d[:, 0] += ds[:, k]
            k += 1
        for (a, b, c, v) in P["nz"]:
            dVin[:, P["in_off"] + a] += v * d[:, c] * G[:, P["sh_off"] + b]
            dG[:, P["sh_off"] + b] += v * d[:, c] * Vin[:, P["in_off"] + a]
    return dVin, dG


class AnalyticAllegro:
    def __init__(self, header, tensors):
        with open(os.path.join(_ROOT, "tables", "allegro_tables.json")) as f:
            tables = json.load(f)
        h = header
        self.L = int(h["l_max"]); self.nl = int(h["num_layers"]); self.B = int(h["num_bessels"])
        self.p = float(h["polynomial_cutoff_p"]); self.S = int(h["num_scalar_features"])
        self.U = int(h["num_tensor_features"]); self.D = int(h["mlp_depth"])
        self.T = int(h["num_types"]); self.inv_sqrt_n = 1.0 / math.sqrt(float(h["avg_num_neighbors"]))
        self.nsh = (self.L + 1) ** 2
        self.kinds = [tables["L"][str(self.L)]["kinds"][k] for k in layer_kinds(self.nl)]
        self.w = {k: np.asarray(v, dtype=np.float64) for k, v in tensors.items()}
        self.lsel = np.array([int(math.isqrt(k)) for k in range(self.nsh)])

    def mlp_fwd(self, prefix, x):
        zs, ds = [], []
        for i in range(self.D + 1):
            z = x @ self.w["%s.w%d" % (prefix, i)]
            if i < self.D:
                x, d = silu(z)
                ds.append(d)
            else:
                x = z
        return x, ds

    def mlp_bwd(self, prefix, ds, dout):
        g = dout
        for i in range(self.D, -1, -1):
            g = g @ self.w["%s.w%d" % (prefix, i)].T
            if i > 0:
                g = g * ds[i - 1]
        return g

    def run(self, rvec, center, zi, zj, n_centers, store_quant=None):
        """rvec [E,3] (x_j - x_i), center [E] centre slots, zi/zj model types.
        returns dict with energies, per-edge gradient g_e = dE_tot/drvec and intermediates.
        store_quant (study hook, tests/study_bf16_storage.py): dict {"zd": f, "v": f, "dv": f} of functions
        applied to the arrays the CUDA pipeline keeps in HBM between kernels -- the activation record
        (act'(z1), act'(z2), m), V^k (k >= 1) and dV^k -- to predict what a narrower storage format costs."""
        sq = store_quant or {}
        qzd = sq.get("zd", lambda t: t)
        qv = sq.get("v", lambda t: t)
        qdv = sq.get("dv", lambda t: t)
        w, L, U, S, nsh = self.w, self.L, self.U, self.S, self.nsh
        E = len(rvec)
        I = {}
        rc = w["cutoff_table"][zi, zj]
        r = np.sqrt((rvec * rvec).sum(1))
        n = rvec / r[:, None]
        xr = r / rc
        p = self.p
        u = np.where(xr < 1, 1 - 0.5 * (p + 1) * (p + 2) * xr ** p + p * (p + 2) * xr ** (p + 1) - 0.5 * p * (p + 1) * xr ** (p + 2), 0.0)
        du_dr = np.where(xr < 1, (-0.5 * p * (p + 1) * (p + 2) * xr ** (p - 1) + p * (p + 1) * (p + 2) * xr ** p
                                  - 0.5 * p * (p + 1) * (p + 2) * xr ** (p + 1)) / rc, 0.0)
        nn = np.arange(1, self.B + 1)[None, :]
        pref = np.sqrt(2.0 / rc)[:, None]
        arg = nn * np.pi * xr[:, None]
        bes = pref * np.sin(arg) / r[:, None]
        dbes = pref * (nn * np.pi / rc[:, None] * np.cos(arg) / r[:, None] - np.sin(arg) / (r * r)[:, None])
        Y, gradY = sh_and_grad(n, L)
        T = self.T
        onehot = np.zeros((E, 2 * T))
        onehot[np.arange(E), zi] = 1; onehot[np.arange(E), T + zj] = 1
        in2b = np.concatenate([onehot, bes * u[:, None]], 1)
        # ---------------- F0
        m0, d2b = self.mlp_fwd("twobody", in2b)
        xs = [m0 * u[:, None]]
        d2b = [qzd(t) for t in d2b]
        ms = [qzd(m0)]
        w0 = (xs[0] @ w["embed_linear"]).reshape(E, L + 1, U)
        V = w0[:, self.lsel, :] * Y[:, :, None]
        Vs = [V]
        Gs, envw, mlpd, ss = [], [], [], []
        for k in range(self.nl):
            K = self.kinds[k]
            wk = (xs[k] @ w["layer%d.env_linear" % k]).reshape(E, L + 1, U)
            envw.append(wk)
            Gam = np.zeros((n_centers, nsh, U))
            np.add.at(Gam, center, wk[:, self.lsel, :] * Y[:, :, None])
            Gam *= self.inv_sqrt_n
            Gs.append(Gam)
            last = k == self.nl - 1
            Vout, s = tp_fwd(K, Vs[k], Gam[center], w["layer%d.omega" % k], not last)
            ss.append(s)
            mk, dk = self.mlp_fwd("layer%d.mlp" % k, np.concatenate([xs[k], s.reshape(E, -1)], 1))
            mlpd.append([qzd(t) for t in dk]); ms.append(qzd(mk))
            al = float(w["layer%d.alpha" % k][0])
            a_, b_ = 1 / math.sqrt(1 + al * al), al / math.sqrt(1 + al * al)
            xs.append(a_ * xs[k] + b_ * mk * u[:, None])
            if not last:
                Vs.append(qv(Vout))
        z = xs[-1] @ w["readout.w0"]
        r1, dr1 = silu(z)
        e_edge = (r1 @ w["readout.w1"])[:, 0]
        I.update(x=xs, V=Vs, Gamma=Gs, e_edge=e_edge, Y=Y, u=u)
        # ---------------- backward.  dE_tot/dE_e:
        ge = self.inv_sqrt_n * w["scales"][zi]
        dx = ((ge[:, None] * w["readout.w1"][:, 0][None, :]) * dr1) @ w["readout.w0"].T
        du = np.zeros(E)
        dY = np.zeros((E, nsh))
        dV = None
        dGs = [None] * self.nl
        dXs = [None] * (self.nl + 1)
        dXs[self.nl] = dx.copy()
        for k in range(self.nl - 1, -1, -1):
            K = self.kinds[k]
            al = float(w["layer%d.alpha" % k][0])
            a_, b_ = 1 / math.sqrt(1 + al * al), al / math.sqrt(1 + al * al)
            # phase 1 (per edge): residual, envelope, MLP backward, TP backward
            dxt = b_ * dx
            du += (dxt * ms[k + 1]).sum(1)
            din = self.mlp_bwd("layer%d.mlp" % k, mlpd[k], dxt * u[:, None])
            dx = a_ * dx + din[:, :S]
            dsk = din[:, S:].reshape(E, K["n0"], U)
            dVin, dG = tp_bwd(K, Vs[k], Gs[k][center], w["layer%d.omega" % k], dV, dsk)
            dGam = np.zeros((n_centers, nsh, U))
            np.add.at(dGam, center, dG)          # segmented sum over the centre's edges  [SYNC]
            dGs[k] = dGam
            # phase 2 (per edge, needs dGamma of the centre): env weights and Y
            dwy = dGam[center] * self.inv_sqrt_n                     # d/d(w*Y) [E,nsh,U]
            dY += (dwy * envw[k][:, self.lsel, :]).sum(2)
            dwk = np.zeros((E, L + 1, U))
            np.add.at(dwk, (slice(None), self.lsel), dwy * Y[:, :, None])
            dx = dx + dwk.reshape(E, -1) @ w["layer%d.env_linear" % k].T
            dV = qdv(dVin) if k > 0 else dVin
            dXs[k] = dx.copy()
        # V0 = w0 (x) Y
        dY += (dV * w0[:, self.lsel, :]).sum(2)
        dw0 = np.zeros((E, L + 1, U))
        np.add.at(dw0, (slice(None), self.lsel), dV * Y[:, :, None])
        dx = dx + dw0.reshape(E, -1) @ w["embed_linear"].T
        # x0 = m0 * u ; two-body MLP
        du += (dx * ms[0]).sum(1)
        din = self.mlp_bwd("twobody", d2b, dx * u[:, None])
        dbu = din[:, 2 * T:]
        dr = (dbu * (dbes * u[:, None] + bes * du_dr[:, None])).sum(1) + du * du_dr
        q = np.einsum("el,elk->ek", dY, gradY)
        g = dr[:, None] * n + (q - n * (n * q).sum(1, keepdims=True)) / r[:, None]
        e_atom = np.zeros(n_centers)
        np.add.at(e_atom, center, e_edge)
        I.update(dX=dXs, dGamma=dGs, dY=dY, du=du, g=g, e_atom_raw=e_atom * self.inv_sqrt_n)
        return I
