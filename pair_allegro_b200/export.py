"""One-time OFFLINE weight exporter: TorchScript `.nequip.pth` -> `.alg` weight file.

This is the only place of the product where libtorch/PyTorch touches model data
(BASELINE.json north_star: "libtorch appears only in a one-time offline weight exporter");
the timestep loop reads the `.alg` file through the C-ABI (include/allegro_b200.h) and never
imports torch.

`.alg` layout (little endian):
    text header, '\\n'-separated `key value...` lines, first line `ALGB200 1`,
    second line `data_offset <bytes>`, terminated by a line `end`;
    `tensor <name> <f32|f64> <ndim> <dims...> <offset> <nbytes>` lines give offsets relative
    to data_offset; blobs are 64-byte aligned.
The five metadata keys the reference reads from the TorchScript archive
(/root/reference/pair_nequip_allegro.cpp:214-220: r_max, per_edge_type_cutoff, type_names,
num_types, allow_tf32) are carried verbatim.
"""
import json
import os
from typing import Dict

import numpy as np

MAGIC = "ALGB200 1"
METADATA_KEYS = ["r_max", "per_edge_type_cutoff", "type_names", "num_types", "allow_tf32"]
CONFIG_KEY = "allegro_b200_config"


def layer_kinds(n_layers: int):
    return {1: ["A"], 2: ["B", "A"], 3: ["C", "D", "A"]}[n_layers]


def write_alg(path: str, header: Dict[str, str], tensors: Dict[str, np.ndarray]) -> None:
    lines = []
    blobs = []
    off = 0
    for name, arr in tensors.items():
        arr = np.ascontiguousarray(arr)
        dt = {"float32": "f32", "float64": "f64"}[str(arr.dtype)]
        off = (off + 63) // 64 * 64
        lines.append("tensor %s %s %d %s %d %d" % (name, dt, arr.ndim, " ".join(str(d) for d in arr.shape), off, arr.nbytes))
        blobs.append((off, arr.tobytes()))
        off += arr.nbytes
    hdr_lines = [MAGIC, "data_offset %010d"] + ["%s %s" % (k, v) for k, v in header.items()] + lines + ["end"]
    text = "\n".join(hdr_lines) + "\n"
    data_offset = (len(text.encode()) + 4095) // 4096 * 4096
    text = text % data_offset
    raw = text.encode()
    assert len(raw) <= data_offset
    with open(path, "wb") as f:
        f.write(raw)
        f.write(b"\0" * (data_offset - len(raw)))
        pos = 0
        for o, b in blobs:
            f.write(b"\0" * (o - pos))
            f.write(b)
            pos = o + len(b)


def export_alg(pth_path: str, alg_path: str) -> Dict[str, str]:
    """Convert a TorchScript Allegro model written by this repo's model definition."""
    import torch  # offline only

    extra = {k: "" for k in METADATA_KEYS + [CONFIG_KEY]}
    m = torch.jit.load(pth_path, map_location="cpu", _extra_files=extra)
    extra = {k: (v.decode() if isinstance(v, bytes) else v) for k, v in extra.items()}
    if not extra[CONFIG_KEY]:
        raise RuntimeError(
            "%s carries no '%s' entry: only models of this repo's network spec (DESIGN.md) can be exported; "
            "importing nequip/allegro checkpoints is a NEXT row (SURVEY.md section 8f-2)" % (pth_path, CONFIG_KEY))
    cfg = json.loads(extra[CONFIG_KEY])
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    L, nl = int(cfg["l_max"]), int(cfg["num_layers"])
    D = int(cfg["mlp_depth"])
    header = {k: extra[k] for k in METADATA_KEYS}
    header.update({
        "l_max": str(L), "num_layers": str(nl), "num_bessels": str(int(cfg["num_bessels"])),
        "polynomial_cutoff_p": repr(float(cfg["polynomial_cutoff_p"])),
        "num_scalar_features": str(int(cfg["num_scalar_features"])),
        "num_tensor_features": str(int(cfg["num_tensor_features"])),
        "mlp_depth": str(D), "mlp_width": str(int(cfg["mlp_width"])),
        "readout_width": str(int(cfg["readout_width"])),
        "avg_num_neighbors": repr(float(cfg["avg_num_neighbors"])),
        "layer_kinds": " ".join(layer_kinds(nl)),
        "model_dtype": "float32",
    })
    f32 = lambda t: t.to(torch.float32).numpy()
    tensors = {}
    for i in range(D + 1):
        tensors["twobody.w%d" % i] = f32(sd["twobody.weights.%d" % i])
    tensors["embed_linear"] = f32(sd["embed_linear"])
    for k in range(nl):
        tensors["layer%d.env_linear" % k] = f32(sd["layers.%d.env_linear" % k])
        tensors["layer%d.omega" % k] = f32(sd["layers.%d.omega" % k])
        for i in range(D + 1):
            tensors["layer%d.mlp.w%d" % (k, i)] = f32(sd["layers.%d.mlp.weights.%d" % (k, i)])
        tensors["layer%d.alpha" % k] = f32(sd["layers.%d.alpha" % k])
    tensors["readout.w0"] = f32(sd["readout.weights.0"])
    tensors["readout.w1"] = f32(sd["readout.weights.1"])
    tensors["scales"] = sd["scales"].to(torch.float64).numpy()
    tensors["shifts"] = sd["shifts"].to(torch.float64).numpy()
    tensors["cutoff_table"] = sd["cutoff_table"].to(torch.float64).numpy()
    write_alg(alg_path, header, tensors)
    return header


def read_alg(path: str):
    """pure-numpy reader (tests / tooling); the product loader is the C++ one in csrc/."""
    with open(path, "rb") as f:
        raw = f.read()
    head = raw[:raw.index(b"\nend\n") + 5].decode()
    lines = head.split("\n")
    assert lines[0] == MAGIC
    data_offset = int(lines[1].split()[1])
    header, tensors = {}, {}
    for ln in lines[2:]:
        if ln == "end" or not ln:
            continue
        key, _, val = ln.partition(" ")
        if key == "tensor":
            t = val.split()
            name, dt, nd = t[0], t[1], int(t[2])
            dims = [int(v) for v in t[3:3 + nd]]
            off, nb = int(t[3 + nd]), int(t[4 + nd])
            tensors[name] = np.frombuffer(raw, dtype={"f32": np.float32, "f64": np.float64}[dt],
                                          count=int(np.prod(dims)) if dims else 1, offset=data_offset + off).reshape(dims)
        else:
            header[key] = val
    return header, tensors


if __name__ == "__main__":
    import sys
    if len(sys.argv) != 3:
        print("usage: python -m pair_allegro_b200.export <model>.nequip.pth <model>.alg")
        sys.exit(2)
    print(export_alg(sys.argv[1], sys.argv[2]))
