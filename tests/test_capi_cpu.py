"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every
symbol include/allegro_b200.h declares; without a GPU it fails LOUDLY (no CPU fallback); the
host-side pair-style mirror keeps the reference's argument/error behaviour."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from helpers import alg_path


def test_header_symbols_exported(ensure_built):
    from pair_allegro_b200 import capi
    hdr = open(os.path.join(ROOT, "include", "allegro_b200.h")).read()
    declared = sorted(set(re.findall(r"ALG_API[^;]*?\b(alg_[a-z_0-9]+)\s*\(", hdr)))
    assert declared, "no prototypes found in the header"
    lib = ctypes.CDLL(ensure_built)
    for sym in declared:
        assert hasattr(lib, sym), "library does not export " + sym
    assert sorted(capi.EXPORTS) == declared
    assert b"sm_100a" in capi.load_library().alg_version()


def test_library_has_sm100a_code(ensure_built):
    """the shipped binary carries sm_100a SASS (not PTX-only, not another arch)"""
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", ensure_built], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_no_cpu_fallback(ensure_built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pair_allegro_b200 import capi
    with pytest.raises(capi.AllegroError) as ei:
        capi.Handle(alg_path("Cu_r5"), 0)
    assert ei.value.code == -3 and "no CPU fallback" in str(ei.value)


def test_neighbor_builder_and_comm_fail_loudly_without_gpu(ensure_built):
    """the helper objects of the C-ABI (device neighbour list, NCCL halo) have no CPU path either"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pair_allegro_b200 import capi
    with pytest.raises(capi.AllegroError):
        capi.NeighborBuilder(0)
    with pytest.raises(capi.AllegroError):
        capi.Comm(0, 1, 0, None)


def test_sass_carries_tensor_core_and_tma_instructions(ensure_built):
    """the hot kernels are what DESIGN.md says they are: tcgen05 MMAs (UTCHMMA), TMEM stores/loads (STTM / LDTM), bulk copies
    (UBLKCP) in the tile kernels, warp-level HMMA in the width-generic GEMM"""
    import subprocess
    sass = subprocess.run(["cuobjdump", "-sass", ensure_built], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "STTM", "LDTM", "UBLKCP", "HMMA"):
        assert mnemonic in sass, mnemonic


def test_create_rejects_bad_file(ensure_built, tmp_path):
    """argument validation happens before any device work only for null args; a malformed file on a
    GPU-less box still reports the missing device first -- both are loud errors"""
    from pair_allegro_b200 import capi
    p = tmp_path / "bad.alg"
    p.write_bytes(b"not a weight file")
    with pytest.raises(capi.AllegroError):
        capi.Handle(str(p), 0)


def test_pair_style_argument_errors():
    """same messages / conditions as pair_nequip_allegro.cpp:171,185-192,205"""
    from pair_allegro_b200.pair import PairAllegroB200
    p = PairAllegroB200()
    with pytest.raises(RuntimeError, match="too many arguments"):
        p.settings(["x"])
    with pytest.raises(RuntimeError, match="Incorrect args for pair coefficients"):
        p.coeff(["*", "*", "m.alg"], 2)                     # missing type names
    with pytest.raises(RuntimeError, match="Incorrect args for pair coefficients"):
        p.coeff(["1", "*", "m.alg", "Cu"], 1)
    with pytest.raises(RuntimeError, match="Only accepts model paths"):
        p.coeff(["*", "*", "model.pt", "Cu"], 1)
    with pytest.raises(RuntimeError, match="requires newton pair on"):
        p.init_style(newton_pair=0)
    with pytest.raises(RuntimeError, match="requires atom IDs"):
        p.init_style(tag_enable=0)


def test_alg_roundtrip_and_metadata():
    from pair_allegro_b200.export import read_alg
    hdr, ten = read_alg(alg_path("Cu2AgO4_r5"))
    assert hdr["type_names"] == "Ag Cu O" and hdr["num_types"] == "3" and hdr["allow_tf32"] == "0"
    assert len(hdr["per_edge_type_cutoff"].split()) == 9
    assert ten["twobody.w0"].shape == (2 * 3 + 8, 64)
    assert ten["layer0.mlp.w0"].shape == (64 + 4 * 32, 64)
    assert ten["scales"].dtype == np.float64


def test_generated_tables_are_current():
    """tp_gen.cuh / allegro_tables.json are what tools/gen_tables.py generates (structure check)"""
    import importlib.util
    import json
    spec = importlib.util.spec_from_file_location("gen_tables", os.path.join(ROOT, "tools", "gen_tables.py"))
    gt = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gt)
    t = gt.build_tables()
    old = json.load(open(os.path.join(ROOT, "tables", "allegro_tables.json")))
    for L in ("1", "2", "3"):
        for k in "ABCD":
            a, b = t["L"][L]["kinds"][k], old["L"][L]["kinds"][k]
            assert a["n_paths"] == b["n_paths"] and a["din"] == b["din"] and a["dout"] == b["dout"]
            for pa, pb in zip(a["paths"], b["paths"]):
                assert [q[:3] for q in pa["nz"]] == [q[:3] for q in pb["nz"]]
                np.testing.assert_allclose([q[3] for q in pa["nz"]], [q[3] for q in pb["nz"]], atol=1e-12)


def test_cpp_pair_style_error_paths_cpu(ensure_built):
    """the C++ pair style (src/pair_allegro_b200.cpp) under the lmpshim harness: LAMMPS-style
    errors surface as exceptions with the reference's messages; without a GPU coeff() fails loudly"""
    import torch
    from helpers import load_golden
    from lmpshim import driver
    if not os.path.exists(driver.OURS_LIB):
        import __graft_entry__ as g
        g.build()
    atom, lst, z = load_golden("Cu_r5")
    lmp = driver.ShimLammps(driver.OURS_LIB, atom, lst)
    with pytest.raises(driver.ShimError, match="too many arguments"):
        lmp.pair_style(["x"])
    with pytest.raises(driver.ShimError, match="Incorrect args for pair coefficients"):
        lmp.pair_coeff(["*", "*", alg_path("Cu_r5")])
    with pytest.raises(driver.ShimError, match="Only accepts model paths"):
        lmp.pair_coeff(["*", "*", "model.pt", "Cu"])
    if not torch.cuda.is_available():
        with pytest.raises(driver.ShimError, match="no CPU fallback"):
            lmp.pair_coeff(["*", "*", alg_path("Cu_r5"), "Cu"])
    assert lmp.flags()["restartinfo"] == 0 and lmp.flags()["manybody_flag"] == 1


def test_header_documents_every_option_key():
    """every key alg_set_option accepts (csrc/alg_api.cu) is documented in include/allegro_b200.h"""
    import re
    src = open(os.path.join(ROOT, "pair_allegro_b200", "csrc", "alg_api.cu")).read()
    body = src[src.index('extern "C" int alg_set_option'):]
    body = body[:body.index("\n}\n")]
    keys = set(re.findall(r'k == "([a-z_]+)"', body))
    hdr = open(os.path.join(ROOT, "include", "allegro_b200.h")).read()
    assert keys >= {"filter", "chunk_edges", "gemm", "precision", "neigh_ago"}
    missing = [k for k in keys if '"%s"' % k not in hdr]
    assert not missing, missing


def test_bench_measured_peaks_parser():
    """bench.py accepts the driver-written MEASURED_PEAKS.json in any plausible nesting / unit"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    r = b.parse_measured_peaks({"hbm_gbs": 6551, "bf16_tflops": 1648, "bf16_tflops_sustained": 1384})
    assert r["bf16_sustained"] == 1384 and r["bf16"] == 1648 and r["hbm_gbs"] == 6551 and r["source"] == "measured"
    r = b.parse_measured_peaks({"hbm": {"copy_GBs_sustained": 6551.0}, "bf16": {"dense_TFLOPs_burst": 1648.2, "dense_TFLOPs_sustained": 1384.1}})
    assert abs(r["bf16_sustained"] - 1384.1) < 1e-9 and abs(r["bf16"] - 1648.2) < 1e-9
    r = b.parse_measured_peaks({"peaks": {"HBM_TB_s": 6.55, "bf16_dense_gflops": 1650000.0}})
    assert abs(r["hbm_gbs"] - 6550.0) < 1e-6 and abs(r["bf16_sustained"] - 1650.0) < 1e-6
    assert b.parse_measured_peaks({"foo": 1}) is None
