"""Random-initialised `.alg` weight files for benchmarking and property tests (numpy only).

There is no network for checkpoints (and no trainer offline), so `bench.py` evaluates
random-init weights of the architecture BASELINE.json names.  Distributions follow SURVEY.md
section 8(d) "Value distributions": every MLP / linear weight ~ N(0,1)/sqrt(fan_in)
(forward-normalised), per-path mixing weights omega ~ N(0,1)/sqrt(#paths into the output irrep),
residual parameter alpha ~ U(0.5, 1.5), per-type scale 1 / shift 0 unless given.
The file is the same format `pair_allegro_b200.export` writes; `oracle.allegro_torch.model_from_alg`
loads it into the torch oracle so that both arms of a comparison evaluate identical weights.
"""
import json
import os
from typing import Dict

import numpy as np

from .export import layer_kinds, write_alg

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def default_config(**kw) -> Dict:
    """Hyper-parameters of the reference's test model (tests/test_data/test_repro_allegro.yaml:86-99)."""
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


def header_from_config(cfg: Dict) -> Dict[str, str]:
    """`.alg` text header: the five metadata keys of pair_nequip_allegro.cpp:214-220 + hyper-parameters."""
    pc = cfg["per_edge_type_cutoff"]
    T = len(cfg["type_names"])
    return {
        "r_max": repr(float(cfg["r_max"])),
        "per_edge_type_cutoff": "" if pc is None else " ".join(repr(float(v)) for v in np.asarray(pc, dtype=np.float64).reshape(-1)),
        "type_names": " ".join(cfg["type_names"]),
        "num_types": str(T),
        "allow_tf32": "1" if cfg["allow_tf32"] else "0",
        "l_max": str(int(cfg["l_max"])), "num_layers": str(int(cfg["num_layers"])),
        "num_bessels": str(int(cfg["num_bessels"])),
        "polynomial_cutoff_p": repr(float(cfg["polynomial_cutoff_p"])),
        "num_scalar_features": str(int(cfg["num_scalar_features"])),
        "num_tensor_features": str(int(cfg["num_tensor_features"])),
        "mlp_depth": str(int(cfg["mlp_depth"])), "mlp_width": str(int(cfg["mlp_width"])),
        "readout_width": str(int(cfg["readout_width"])),
        "avg_num_neighbors": repr(float(cfg["avg_num_neighbors"])),
        "layer_kinds": " ".join(layer_kinds(int(cfg["num_layers"]))),
        "model_dtype": "float32",
    }


def random_tensors(cfg: Dict) -> Dict[str, np.ndarray]:
    with open(os.path.join(_ROOT, "tables", "allegro_tables.json")) as f:
        tables = json.load(f)
    rng = np.random.default_rng(int(cfg["seed"]))
    L, nl = int(cfg["l_max"]), int(cfg["num_layers"])
    T, B = len(cfg["type_names"]), int(cfg["num_bessels"])
    S, U = int(cfg["num_scalar_features"]), int(cfg["num_tensor_features"])
    D, H, R = int(cfg["mlp_depth"]), int(cfg["mlp_width"]), int(cfg["readout_width"])
    kinds = tables["L"][str(L)]["kinds"]

    def lin(n_in, n_out):
        return (rng.standard_normal((n_in, n_out)) / np.sqrt(n_in)).astype(np.float32)

    def mlp(prefix, dims, out):
        for i in range(len(dims) - 1):
            out["%s%d" % (prefix, i)] = lin(dims[i], dims[i + 1])

    t: Dict[str, np.ndarray] = {}
    mlp("twobody.w", [2 * T + B] + [H] * D + [S], t)
    t["embed_linear"] = lin(S, (L + 1) * U)
    for k, kind_name in enumerate(layer_kinds(nl)):
        kind = kinds[kind_name]
        paths = kind["paths"]
        n0 = sum(1 for p in paths if p["scalar"])
        fan = [0] * len(kind["out_irreps"])
        for p in paths:
            fan[p["o3"]] += 1
        t["layer%d.env_linear" % k] = lin(S, (L + 1) * U)
        om = rng.standard_normal((len(paths), U))
        for ip, p in enumerate(paths):
            om[ip] /= np.sqrt(fan[p["o3"]])
        t["layer%d.omega" % k] = om.astype(np.float32)
        mlp("layer%d.mlp.w" % k, [S + U * n0] + [H] * D + [S], t)
        t["layer%d.alpha" % k] = (0.5 + rng.random(1)).astype(np.float32)
    t["readout.w0"] = lin(S, R)
    t["readout.w1"] = lin(R, 1)
    t["scales"] = np.asarray(cfg["per_type_energy_scales"], dtype=np.float64)
    t["shifts"] = np.asarray(cfg["per_type_energy_shifts"], dtype=np.float64)
    pc = cfg["per_edge_type_cutoff"]
    t["cutoff_table"] = (np.full((T, T), float(cfg["r_max"])) if pc is None
                         else np.asarray(pc, dtype=np.float64).reshape(T, T)).astype(np.float64)
    return t


def random_alg(cfg: Dict, path: str) -> Dict[str, str]:
    """write a random-init model of configuration `cfg` (see default_config) to `path` (.alg)"""
    header = header_from_config(cfg)
    write_alg(path, header, random_tensors(cfg))
    return header
