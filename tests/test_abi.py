"""The C-ABI library: builds, loads, exports every symbol include/satsuma_xcorr.h declares, struct
layouts match the reference's wire structs, and the product path fails loudly without a GPU
(no CPU fallback).  CPU only -- no compute calls."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "satsuma_xcorr.h")


def declared_symbols():
    """every function declared in include/*.h (satsuma_xcorr.h: the hot path; satsuma_kmatch.h: k-mer seeding)"""
    syms = set()
    for name in sorted(os.listdir(os.path.join(ROOT, "include"))):
        if not name.endswith(".h"):
            continue
        text = open(os.path.join(ROOT, "include", name)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        syms |= set(re.findall(r"\b(sx_[a-z_0-9]+)\s*\(", text))
    return sorted(syms)


def test_header_declares_the_boundary():
    syms = declared_symbols()
    for must in ("sx_create", "sx_destroy", "sx_set_targets", "sx_set_queries", "sx_align_blocks",
                 "sx_align_pairs", "sx_tap_xcorr", "sx_tap_candidates", "sx_tap_segments", "sx_tap_signal",
                 "sx_last_error", "sx_get_stats", "sx_build_prob_table", "sx_set_prob_table", "sx_multi_create",
                 "sx_multi_align_blocks", "sx_kmatch", "sx_device_count"):
        assert must in syms


def test_library_exports_every_declared_symbol(sx):
    lib = C.CDLL(sx.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/satsuma_xcorr.h but not exported"


def test_wire_struct_layouts(sx):
    # t_result = 72 bytes, t_pair = 28 bytes (analysis/WorkQueue.h:17-33)
    assert sx.RESULT_DTYPE.itemsize == 72
    assert sx.PAIR_DTYPE.itemsize == 28
    assert sx.RESULT_DTYPE.fields["prob"][1] == 56 and sx.RESULT_DTYPE.fields["reverse"][1] == 48
    assert sx.PAIR_DTYPE.fields["fast"][1] == 16 and sx.PAIR_DTYPE.fields["slave_id"][1] == 20
    assert sx.PAIR_DTYPE.fields["status"][1] == 24


def test_default_config_is_slave_semantics(sx):
    cfg = sx.default_config()
    assert (cfg.t_chunk, cfg.q_chunk) == (4096, 4096)
    assert (cfg.cutoff, cfg.cutoff_fast) == (1.8, 2.9)
    assert cfg.min_prob == 0.99  # Slave.cc:76, SURVEY Q9
    assert cfg.min_len == 0 and cfg.use_prob_table == 0 and cfg.rc_coord_mode == 0


def test_bad_config_rejected(sx):
    lib = sx.load_library()
    for kw in (dict(t_chunk=5000), dict(t_chunk=512), dict(q_chunk=0), dict(q_chunk=8193), dict(abi_version=99)):
        cfg = sx.default_config(**kw)
        h = C.c_void_p()
        rc = lib.sx_create(C.byref(cfg), C.byref(h))
        assert rc == sx.SX_ERR_ARG, kw
        assert lib.sx_last_error()


def test_no_cpu_fallback(sx):
    """Without a CUDA device the product refuses to run instead of silently computing on the CPU."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(sx.SatsumaError) as ei:
        sx.XCorrEngine()
    assert ei.value.code == sx.SX_ERR_CUDA
    assert "no CPU fallback" in str(ei.value)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "satsuma2_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "sx_oracle" not in text and "liboracle" not in text and "libsatsuma_ref" not in text, f


def test_prob_table_builder_matches_oracle(sx, oracle_lib):
    """Host-side ProbTable::Setup in the library (libm) == oracle == reference (golden rows)."""
    tab = sx.build_prob_table(800001.0)
    exp = oracle_lib.prob_table(800001.0)
    assert np.array_equal(tab[1:], exp[1:])


def test_committed_bench_lines_follow_the_contract():
    """The bench lines kept under profiles/ carry every key the measurement contract names."""
    import glob
    import json

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    files = sorted(glob.glob(os.path.join(root, "profiles", "r[12]_bench_*.json")))
    assert files
    for f in files:
        d = json.load(open(f))
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                  "vs_baseline", "dtype", "data", "config", "e2e"):
            assert k in d, (f, k)
        assert d["metric"] == "chunk_pair_xcorrs_per_sec" and d["unit"] == "chunk-pairs/s" and d["warmup"] >= 3
        assert "workload" in d["config"] and d["vs_baseline"] is None
        assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"])
        if d.get("impl") == "reference":
            assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["e2e"]["h2d_bytes_per_step"] == 0
            continue
        assert d["gpu_launches"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0 and "clocks" in d
        if "roofline" in d:
            r = d["roofline"]
            assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r)
            assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
