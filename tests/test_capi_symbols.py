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


def test_newer_entry_points_validate_before_touching_the_device():
    """Error behaviour at the boundary (negative status + thread-local message), no GPU involved."""
    import ctypes as C
    from pygda_b200 import _lib
    lib = _lib.load()
    out = C.c_void_p(0)
    cases = [
        (lib.gda_spmm_f32(None, 0, None, 4, None, 4, 4, None, 0, 0.0, 0, None, None, 0, None), b"graph is NULL"),
        (lib.gda_spmm_nb_f32(None, 0, 2, None, 4, 0, None, 4, 0, 4, None, 0, 0.0, 0, None, None, 0, None), b"graph is NULL"),
        (lib.gda_bias_act_dropout_rep_fwd(None, None, None, -1, 4, 2, 0, 0.0, 0, None, None), b"negative"),
        (lib.gda_bias_act_dropout_rep_bwd(None, None, None, None, None, 4, 4, 3, 0, 0, 0.0, 0, None, None), b"rep must be 1 or 2"),
        (lib.gda_ppmi_create(None, 0, 10, 0, 40, 0, None, C.byref(out)), b"path_len"),
        (lib.gda_ppmi_create(None, 5, 10, 5, 40, 0, None, C.byref(out)), b"edge_index is NULL"),
        (lib.gda_argmax_confusion(None, 10, 65, 65, None, None, None, None, None), b"C <= 64"),
        (lib.gda_unpack_rows_f32(None, None, 3, None, 4, 4, None, 4, None), b"uint16 or int32"),
        (lib.gda_collate_graphs(None, 0, None, 0, None, None, None, 1, None, None, 0, 0, None, None, None, None), b"bad size"),
        (lib.gda_gemm_bf16(0, 0, -1, 64, 64, None, 64, None, 64, None, 64, 1, None, 0, None), b"negative"),
        (lib.gda_attention2_bwd(None, None, 4, 9000, None, None, None, None, None, None, None), b"H <= 8192"),
        (lib.gda_bern_axpy_f32(None, None, -1, 0.0, 1.0, None, None), b"negative"),
    ]
    for rc, needle in cases:
        assert rc < 0
    # the message is the one of the LAST failing call on this thread
    rc = lib.gda_unpack_rows_f32(None, None, 3, None, 4, 4, None, 4, None)
    assert rc < 0 and b"uint16 or int32" in lib.gda_last_error()
    rc = lib.gda_ppmi_create(None, 0, 10, 0, 40, 0, None, C.byref(out))
    assert rc < 0 and b"path_len" in lib.gda_last_error() and not out.value
    # sizes that mean "nothing to do" succeed without a device
    assert lib.gda_bias_act_dropout_rep_fwd(None, None, None, 0, 4, 2, 0, 0.0, 0, None, None) == 0
    assert lib.gda_gemm_bf16(0, 0, 0, 64, 64, None, 64, None, 64, None, 64, 1, None, 0, None) == 0
    assert lib.gda_wedges_size(None) == -1


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
