"""libgda.so loads on a machine without a GPU and exports every symbol that
include/gda.h declares (no compute calls here)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "gda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gda_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for must in ("gda_graph_create", "gda_spmm_f32", "gda_spmm_bf16", "gda_gemm_f32", "gda_mmd_fwd",
                 "gda_mmd_bwd", "gda_softmax_ce_fwd_bwd", "gda_adam_step", "gda_segment_mean_fwd"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    from pygda_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        pytest.fail("libgda.so missing: run __graft_entry__.build()")
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, f"declared in gda.h but not exported: {missing}"


def test_binding_table_matches_header():
    from pygda_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    lib = _lib.load()
    assert lib.gda_version() == 100 and lib.gda_sm_arch() == 100


def test_argument_errors_do_not_need_a_gpu():
    from pygda_b200 import _lib
    lib = _lib.load()
    rc = lib.gda_gemm_f32(0, 0, -1, 4, 4, 1.0, None, 4, None, 4, 0.0, None, 4, None, 0, None)
    assert rc == -1 and b"negative" in lib.gda_last_error()


def test_product_path_refuses_cpu_tensors():
    import torch
    from pygda_b200 import ops
    with pytest.raises(ValueError, match="GPU only"):
        ops.gemm(torch.zeros(2, 2), torch.zeros(2, 2))


def test_no_oracle_import_in_product():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pygda_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, f"product code imports the oracle: {bad}"
