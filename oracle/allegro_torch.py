"""ORACLE (test infrastructure, never on the product path).

Pure-PyTorch restatement of the Allegro network that the reference executes through
libtorch (`torchscript_model.forward(input_vector)`, /root/reference/pair_nequip_allegro.cpp:425).
The `allegro`/`nequip` Python packages that build the real graph are NOT vendored in the
reference tree and are not installable offline, so PARITY IS UNPINNED
in the golden-vector sense: this module *defines* the network (DESIGN.md "Network spec",
distilled from SURVEY.md Appendix A and the hyper-parameter names of
/root/reference/tests/test_data/test_repro_allegro.yaml:79-103) and both the TorchScript
model the reference glue loads and the CUDA kernels implement it.

I/O contract = the reference's (pair_nequip_allegro.cpp:242-247, 358-392, 524-529, 638-641):
  in : pos f64[Ntot,3], edge_index i64[2,E] (row0 centre, row1 neighbour), atom_types i64[Ntot]
  out: atomic_energy f64[Ntot,1], forces f64[Ntot,3], virial f64[1,3,3]
Metadata keys written into the TorchScript `_extra_files` = pair_nequip_allegro.cpp:214-220.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs import this.
"""
import json
import math
import os
from typing import Dict, List, Optional

import numpy as np
import torch

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ACT_C = 1.6765324703310909  # 1/sqrt(E_{z~N(0,1)} silu(z)^2): keeps activations O(1)

METADATA_KEYS = ["r_max", "per_edge_type_cutoff", "type_names", "num_types", "allow_tf32"]
CONFIG_KEY = "allegro_b200_config"


def load_tables():
    with open(os.path.join(_ROOT, "tables", "allegro_tables.json")) as f:
        return json.load(f)


def layer_kinds(n_layers: int) -> List[str]:
    return {1: ["A"], 2: ["B", "A"], 3: ["C", "D", "A"]}[n_layers]


def default_config(**kw):
    """Hyper-parameters of the reference's test model (test_repro_allegro.yaml:86-99)."""
    cfg = dict(type_names=["Cu", "Pd"], r_max=5.0, per_edge_type_cutoff=None, num_bessels=8,
               polynomial_cutoff_p=6, l_max=2, num_layers=3, num_scalar_features=64,
               num_tensor_features=32, mlp_depth=2, mlp_width=64, readout_width=32,
               avg_num_neighbors=42.0, per_type_energy_scales=None, per_type_energy_shifts=None,
               allow_tf32=False, seed=1)
    cfg.update(kw)
    T = len(cfg["type_names"])
    if cfg["per_type_energy_scales"] is None:
        cfg["per_type_energy_scales"] = [1.0] * T
    if cfg["per_type_energy_shifts"] is None:
        cfg["per_type_energy_shifts"] = [0.0] * T
    return cfg


def real_sh_torch(n: torch.Tensor, lmax: int) -> torch.Tensor:
    """Real spherical harmonics, component normalisation, m=-l..l (same definition as
    tools/gen_tables.py:real_sh).  n: [E,3] unit vectors -> [E,(lmax+1)^2]."""
    x, y, z = n[:, 0], n[:, 1], n[:, 2]
    out = [torch.ones_like(x)]
    if lmax >= 1:
        s3 = math.sqrt(3.0)
        out += [s3 * y, s3 * z, s3 * x]
    if lmax >= 2:
        s15 = math.sqrt(15.0)
        s5 = math.sqrt(5.0)
        out += [s15 * x * y, s15 * y * z, 0.5 * s5 * (3.0 * z * z - 1.0), s15 * x * z,
                0.5 * s15 * (x * x - y * y)]
    if lmax >= 3:
        a = math.sqrt(35.0 / 8.0)
        b = math.sqrt(105.0)
        c = math.sqrt(21.0 / 8.0)
        d = 0.5 * math.sqrt(7.0)
        out += [a * y * (3.0 * x * x - y * y), b * x * y * z, c * y * (5.0 * z * z - 1.0),
                d * (5.0 * z * z * z - 3.0 * z), c * x * (5.0 * z * z - 1.0),
                0.5 * b * (x * x - y * y) * z, a * x * (x * x - 3.0 * y * y)]
    return torch.stack(out, dim=1)


class MLP(torch.nn.Module):
    """bias-free MLP, hidden activation ACT_C*silu, weights stored [in,out]."""

    def __init__(self, dims: List[int], gen: torch.Generator):
        super().__init__()
        self.weights = torch.nn.ParameterList(
            [torch.nn.Parameter(torch.randn(dims[i], dims[i + 1], generator=gen) / math.sqrt(dims[i]))
             for i in range(len(dims) - 1)])
        self.n = len(dims) - 1
        self.act_c = ACT_C

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        i = 0
        for w in self.weights:
            x = x @ w
            if i < self.n - 1:
                x = self.act_c * torch.nn.functional.silu(x)
            i += 1
        return x


class AllegroLayer(torch.nn.Module):
    def __init__(self, kind: Dict, nsh: int, L: int, S: int, U: int, depth: int, width: int,
                 last: bool, gen: torch.Generator):
        super().__init__()
        self.last = last
        self.U = U
        self.nsh = nsh
        self.L = L
        din, dout, paths = kind["din"], kind["dout"], kind["paths"]
        self.din = din
        self.dout = dout
        Q = sum(2 * p["l3"] + 1 for p in paths)
        cbig = torch.zeros(din, nsh, Q, dtype=torch.float64)
        qpath = torch.zeros(Q, dtype=torch.long)
        mixmat = torch.zeros(Q, dout)
        scal_q: List[int] = []
        fan = [0] * len(kind["out_irreps"])
        for p in paths:
            fan[p["o3"]] += 1
        q = 0
        for ip, p in enumerate(paths):
            for (a, b, c, v) in p["nz"]:
                cbig[p["in_off"] + a, p["sh_off"] + b, q + c] = v
            for c in range(2 * p["l3"] + 1):
                qpath[q + c] = ip
                mixmat[q + c, p["out_off"] + c] = 1.0
            if p["scalar"]:
                scal_q.append(q)
            q += 2 * p["l3"] + 1
        self.register_buffer("cbig", cbig)   # float64 master copy; cast to the model dtype in forward
        self.register_buffer("qpath", qpath)
        self.register_buffer("mixmat", mixmat)
        self.register_buffer("scal_q", torch.tensor(scal_q, dtype=torch.long))
        self.n0 = len(scal_q)
        # env-weight linear S -> U*(L+1), columns ordered [l][u]
        self.env_linear = torch.nn.Parameter(torch.randn(S, (L + 1) * U, generator=gen) / math.sqrt(S))
        # per-path per-channel mixing weights omega[path,u] (Variant B); unused for the last layer
        om = torch.randn(len(paths), U, generator=gen)
        for ip, p in enumerate(paths):
            om[ip] /= math.sqrt(fan[p["o3"]])
        self.omega = torch.nn.Parameter(om)
        self.mlp = MLP([S + U * self.n0] + [width] * depth + [S], gen)
        self.alpha = torch.nn.Parameter(0.5 + torch.rand(1, generator=gen))

    def forward(self, x: torch.Tensor, V: torch.Tensor, Y: torch.Tensor, u: torch.Tensor,
                center: torch.Tensor, n_atoms: int, inv_sqrt_n: float):
        E = x.shape[0]
        # lsel[k] = l of SH component k
        w = (x @ self.env_linear).view(E, self.L + 1, self.U)          # [E, l, u]
        lsel = torch.floor(torch.sqrt(torch.arange(self.nsh, device=x.device).to(torch.float32) + 0.5)).to(torch.long)
        wy = w[:, lsel, :] * Y[:, :, None]                              # [E, nsh, U]
        gamma = torch.zeros(n_atoms, self.nsh, self.U, dtype=x.dtype, device=x.device)
        gamma = gamma.index_add(0, center, wy) * inv_sqrt_n             # env sum per centre
        G = gamma[center]                                               # [E, nsh, U]
        T = torch.einsum("eau,ebu,abq->equ", V, G, self.cbig.to(V.dtype))   # [E, Q, U]
        s = T[:, self.scal_q, :]                                        # [E, n0, U]
        sflat = s.reshape(E, self.n0 * self.U)                          # column = k*U + u
        xt = self.mlp(torch.cat([x, sflat], dim=1)) * u[:, None]
        a = self.alpha
        xn = (x + a * xt) / torch.sqrt(1.0 + a * a)
        if self.last:
            return xn, V
        Vn = torch.einsum("equ,qu,qo->eou", T, self.omega[self.qpath], self.mixmat)   # [E, dout, U]
        return xn, Vn


class AllegroOracle(torch.nn.Module):
    """forward(Dict[str,Tensor]) -> Dict[str,Tensor]: the graph the reference glue calls."""

    def __init__(self, cfg: Dict):
        super().__init__()
        tables = load_tables()
        L = int(cfg["l_max"])
        TL = tables["L"][str(L)]
        self.L = L
        self.nsh = int(TL["nsh"])
        self.T = len(cfg["type_names"])
        self.B = int(cfg["num_bessels"])
        self.p = float(cfg["polynomial_cutoff_p"])
        self.r_max = float(cfg["r_max"])
        self.S = int(cfg["num_scalar_features"])
        self.U = int(cfg["num_tensor_features"])
        self.n_layers = int(cfg["num_layers"])
        self.inv_sqrt_n = 1.0 / math.sqrt(float(cfg["avg_num_neighbors"]))
        gen = torch.Generator().manual_seed(int(cfg["seed"]))
        pc = cfg["per_edge_type_cutoff"]
        if pc is None:
            cut = torch.full((self.T, self.T), self.r_max, dtype=torch.float64)
        else:
            cut = torch.tensor(pc, dtype=torch.float64).view(self.T, self.T)
        self.register_buffer("cutoff_table", cut)
        self.register_buffer("bessel_n", torch.arange(1, self.B + 1, dtype=torch.float32))
        D, H = int(cfg["mlp_depth"]), int(cfg["mlp_width"])
        self.twobody = MLP([2 * self.T + self.B] + [H] * D + [self.S], gen)
        self.embed_linear = torch.nn.Parameter(
            torch.randn(self.S, (L + 1) * self.U, generator=gen) / math.sqrt(self.S))
        kinds = layer_kinds(self.n_layers)
        self.layers = torch.nn.ModuleList(
            [AllegroLayer(TL["kinds"][k], self.nsh, L, self.S, self.U, D, H, i == self.n_layers - 1, gen)
             for i, k in enumerate(kinds)])
        self.readout = MLP([self.S, int(cfg["readout_width"]), 1], gen)
        self.register_buffer("scales", torch.tensor(cfg["per_type_energy_scales"], dtype=torch.float64))
        self.register_buffer("shifts", torch.tensor(cfg["per_type_energy_shifts"], dtype=torch.float64))

    def edge_energy(self, rvec: torch.Tensor, center: torch.Tensor, zi: torch.Tensor,
                    zj: torch.Tensor, n_atoms: int) -> torch.Tensor:
        """rvec f64[E,3] (grad leaf) -> per-edge energies (model dtype)."""
        dt = self.embed_linear.dtype
        rc = self.cutoff_table[zi, zj].to(dt)
        rv = rvec.to(dt)
        r = torch.sqrt((rv * rv).sum(dim=1))
        xr = r / rc
        p = self.p
        poly = 1.0 - 0.5 * (p + 1.0) * (p + 2.0) * torch.pow(xr, p) + p * (p + 2.0) * torch.pow(xr, p + 1.0) \
            - 0.5 * p * (p + 1.0) * torch.pow(xr, p + 2.0)
        u = torch.where(xr < 1.0, poly, torch.zeros_like(poly))
        bes = torch.sqrt(2.0 / rc)[:, None] * torch.sin(self.bessel_n.to(dt)[None, :] * (math.pi * xr)[:, None]) / r[:, None]
        onehot_i = torch.nn.functional.one_hot(zi, self.T).to(dt)
        onehot_j = torch.nn.functional.one_hot(zj, self.T).to(dt)
        x = self.twobody(torch.cat([onehot_i, onehot_j, bes * u[:, None]], dim=1)) * u[:, None]
        Y = real_sh_torch(rv / r[:, None], self.L)                       # [E, nsh]
        E = x.shape[0]
        w0 = (x @ self.embed_linear).view(E, self.L + 1, self.U)
        lsel = torch.floor(torch.sqrt(torch.arange(self.nsh, device=x.device).to(torch.float32) + 0.5)).to(torch.long)
        V = w0[:, lsel, :] * Y[:, :, None]                               # [E, nsh, U]
        for layer in self.layers:
            x, V = layer(x, V, Y, u, center, n_atoms, self.inv_sqrt_n)
        return self.readout(x)[:, 0]

    def forward(self, data: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        pos = data["pos"]
        edge_index = data["edge_index"]
        types = data["atom_types"]
        center, nbr = edge_index[0], edge_index[1]
        n_atoms = pos.shape[0]
        rvec = (pos[nbr] - pos[center]).detach().requires_grad_(True)   # f64, r_ij = x_j - x_i
        zi, zj = types[center], types[nbr]
        e_edge = self.edge_energy(rvec, center, zi, zj, n_atoms)
        e_atom = torch.zeros(n_atoms, dtype=torch.float64, device=pos.device)
        e_atom = e_atom.index_add(0, center, e_edge.to(torch.float64)) * self.inv_sqrt_n
        e_atom = e_atom * self.scales[types] + self.shifts[types]
        grads = torch.autograd.grad([e_atom.sum()], [rvec])
        g_opt = grads[0]
        if g_opt is None:
            g = torch.zeros_like(rvec)
        else:
            g = g_opt
        forces = torch.zeros(n_atoms, 3, dtype=torch.float64, device=pos.device)
        forces = forces.index_add(0, center, g)
        forces = forces.index_add(0, nbr, -g)
        w = -torch.einsum("ea,eb->ab", rvec.detach(), g)
        virial = (0.5 * (w + w.t())).unsqueeze(0)
        out: Dict[str, torch.Tensor] = {}
        out["atomic_energy"] = e_atom.detach().unsqueeze(1)
        out["forces"] = forces
        out["virial"] = virial
        out["edge_energy"] = e_edge.detach().to(torch.float64)
        return out


def metadata_from_config(cfg: Dict) -> Dict[str, str]:
    T = len(cfg["type_names"])
    pc = cfg["per_edge_type_cutoff"]
    return {
        "r_max": repr(float(cfg["r_max"])),
        "per_edge_type_cutoff": "" if pc is None else " ".join(repr(float(v)) for v in sum([list(r) for r in pc], [])),
        "type_names": " ".join(cfg["type_names"]),
        "num_types": str(T),
        "allow_tf32": "1" if cfg["allow_tf32"] else "0",
        CONFIG_KEY: json.dumps(cfg),
    }


def build_model(cfg: Dict, dtype=torch.float32) -> AllegroOracle:
    m = AllegroOracle(cfg)
    if dtype == torch.float64:
        # fp64-parameter ground truth: same fp32-representable weights, math in double
        for p in m.parameters():
            p.data = p.data.to(torch.float64)
        for layer in m.layers:
            layer.mixmat = layer.mixmat.to(torch.float64)
    m.eval()
    for p in m.parameters():
        p.requires_grad_(False)
    return m


def save_torchscript(cfg: Dict, path: str, dtype=torch.float32) -> AllegroOracle:
    """Write `<path>` (.nequip.pth) the way nequip-compile would: scripted module +
    metadata in _extra_files (keys read at pair_nequip_allegro.cpp:214-222)."""
    m = build_model(cfg, dtype)
    sm = torch.jit.script(m)
    torch.jit.save(sm, path, _extra_files=metadata_from_config(cfg))
    return m


def config_from_alg_header(hdr: Dict[str, str]) -> Dict:
    """inverse of the `.alg` header written by pair_allegro_b200 (export.py / modelgen.py)"""
    names = hdr["type_names"].split()
    T = len(names)
    pc = hdr.get("per_edge_type_cutoff", "").split()
    return default_config(
        type_names=names, r_max=float(hdr["r_max"]),
        per_edge_type_cutoff=None if not pc else [[float(pc[i * T + j]) for j in range(T)] for i in range(T)],
        num_bessels=int(hdr["num_bessels"]), polynomial_cutoff_p=float(hdr["polynomial_cutoff_p"]),
        l_max=int(hdr["l_max"]), num_layers=int(hdr["num_layers"]),
        num_scalar_features=int(hdr["num_scalar_features"]), num_tensor_features=int(hdr["num_tensor_features"]),
        mlp_depth=int(hdr["mlp_depth"]), mlp_width=int(hdr["mlp_width"]), readout_width=int(hdr["readout_width"]),
        avg_num_neighbors=float(hdr["avg_num_neighbors"]), allow_tf32=hdr.get("allow_tf32", "0") == "1", seed=0)


def model_from_alg(alg_path: str, dtype=torch.float32):
    """Oracle model carrying exactly the weights of a `.alg` file (e.g. one written by
    pair_allegro_b200.modelgen): both arms of a comparison then evaluate identical parameters.
    Returns (model, cfg)."""
    from pair_allegro_b200.export import read_alg
    hdr, ten = read_alg(alg_path)
    cfg = config_from_alg_header(hdr)
    cfg["per_type_energy_scales"] = [float(v) for v in ten["scales"]]
    cfg["per_type_energy_shifts"] = [float(v) for v in ten["shifts"]]
    m = AllegroOracle(cfg)
    D = int(cfg["mlp_depth"])
    sd = {}
    for i in range(D + 1):
        sd["twobody.weights.%d" % i] = ten["twobody.w%d" % i]
    sd["embed_linear"] = ten["embed_linear"]
    for k in range(int(cfg["num_layers"])):
        sd["layers.%d.env_linear" % k] = ten["layer%d.env_linear" % k]
        sd["layers.%d.omega" % k] = ten["layer%d.omega" % k]
        for i in range(D + 1):
            sd["layers.%d.mlp.weights.%d" % (k, i)] = ten["layer%d.mlp.w%d" % (k, i)]
        sd["layers.%d.alpha" % k] = ten["layer%d.alpha" % k]
    sd["readout.weights.0"] = ten["readout.w0"]
    sd["readout.weights.1"] = ten["readout.w1"]
    sd["cutoff_table"] = ten["cutoff_table"]
    missing, unexpected = m.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd.items()}, strict=False)
    assert not unexpected, unexpected
    assert all(("cbig" in k or "qpath" in k or "mixmat" in k or "scal_q" in k or k in ("bessel_n", "scales", "shifts")) for k in missing), missing
    if dtype == torch.float64:
        for p in m.parameters():
            p.data = p.data.to(torch.float64)
        for layer in m.layers:
            layer.mixmat = layer.mixmat.to(torch.float64)
    m.eval()
    for p in m.parameters():
        p.requires_grad_(False)
    return m, cfg


def save_torchscript_from_alg(alg_path: str, pth_path: str, dtype=torch.float32):
    """`.alg` -> `.nequip.pth` the reference glue can load (same weights, metadata in _extra_files)"""
    m, cfg = model_from_alg(alg_path, dtype)
    torch.jit.save(torch.jit.script(m), pth_path, _extra_files=metadata_from_config(cfg))
    return m
