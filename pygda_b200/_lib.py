"""ctypes binding of libgda.so (the C ABI declared in include/gda.h).

There is deliberately no fallback: if the library is missing or a call fails the
caller gets an exception -- the product path never routes through PyTorch
reference ops or the CPU oracle.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GDA_LIB_PATH") or os.path.join(_HERE, "libgda.so")   # override: kernel-variant experiments


class GdaError(RuntimeError):
    pass


_lib = None

i64, i32, f32, u64, vp = C.c_int64, C.c_int, C.c_float, C.c_uint64, C.c_void_p

# name -> (restype, argtypes); every symbol of include/gda.h
SIGNATURES = {
    "gda_version": (i32, []),
    "gda_sm_arch": (i32, []),
    "gda_last_error": (C.c_char_p, []),
    "gda_launch_count": (u64, []),
    "gda_graph_create": (i32, [vp, i64, i64, vp, i32, vp, C.POINTER(vp)]),
    "gda_graph_destroy": (i32, [vp]),
    "gda_graph_info": (i32, [vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)]),
    "gda_graph_export_coo": (i32, [vp, vp, vp, vp]),
    "gda_graph_export_csr": (i32, [vp, i32, vp, vp, vp, vp]),
    "gda_spmm_workspace_bytes": (i64, [vp, i32, i32]),
    "gda_spmm_f32": (i32, [vp, i32, vp, i64, vp, i64, i32, vp, i32, f32, u64, vp, vp, i64, vp]),
    "gda_spmm_k_f32": (i32, [vp, i32, i32, vp, i64, vp, i64, vp, vp, i32, vp, i32, f32, u64, vp, vp, i64, vp]),
    "gda_spmm_nb_f32": (i32, [vp, i32, i32, vp, i64, i64, vp, i64, i64, i32, vp, i32, f32, u64, vp, vp, i64, vp]),
    "gda_graph_unit_weights": (i32, [vp, i32, i32, i32]),
    "gda_row_scale_f32": (i32, [vp, i32, vp, i64, i64, vp, i64, i64, i32, vp]),
    "gda_graph_export_dinv": (i32, [vp, vp, vp]),
    "gda_graph_set_unit_dinv": (i32, [vp, vp, vp]),
    "gda_row_scale_rows_f32": (i32, [vp, i64, i32, vp, i64, i64, vp, i64, i64, i32, vp]),
    "gda_spmm_unw_nb_f32": (i32, [vp, i32, i32, vp, i64, i64, vp, i64, i64, i32, i32, vp, i32, f32, u64, vp, vp, i64, vp]),
    "gda_spmm_k_nb_f32": (i32, [vp, i32, i32, i32, vp, i64, i64, vp, i64, i64, vp, vp, i32, vp, i32, f32, u64, vp, vp,
                                i64, vp]),
    "gda_spmm_peer_k_f32": (i32, [vp, i32, i32, vp, vp, vp, i32, i32, vp, i32, vp, i32, f32, u64, vp, vp, i64,
                                  vp, u64, vp, vp]),
    "gda_spmm_peer_k_dev_f32": (i32, [vp, i32, i32, vp, vp, vp, i32, i32, vp, i32, vp, i32, f32, u64, vp, vp, i64,
                                      vp, vp, vp, vp]),
    "gda_peer_barrier_dev": (i32, [vp, i32, i32, vp, vp, vp]),
    "gda_spmm_halo_f32": (i32, [vp, vp, i64, vp, i64, i32, vp, vp, vp, i32, vp, i32, f32, u64, vp, vp, i64, vp]),
    "gda_spmm_unw_halo_f32": (i32, [vp, i32, vp, i64, i64, vp, i64, i64, i32, vp, vp, vp, vp, i32, vp, i64, vp]),
    "gda_push_rows_f32": (i32, [vp, i64, vp, vp, vp, i64, vp, i32, i64, i32, vp]),
    "gda_spmm_push_f32": (i32, [vp, i32, vp, vp, i32, i32, i64, i64, i32, vp, i32, f32, u64, vp, vp, i64, vp]),
    "gda_spmm_push_k_f32": (i32, [vp, i32, i32, vp, vp, vp, i32, i32, vp, i32, vp, i32, f32, u64, vp, vp, i64, vp, vp, vp,
                                  vp]),
    "gda_spmm_bf16": (i32, [vp, i32, vp, i64, vp, i64, i32, vp, i32, f32, u64, vp, vp, i64, vp]),
    "gda_graph_partition": (i32, [vp, i64, i64, i64, vp, C.POINTER(vp)]),
    "gda_spmm_peer_f32": (i32, [vp, i32, vp, i32, i32, i64, vp, i64, i32, vp, i32, f32, u64, vp, vp, i64, vp]),
    "gda_sym_alloc": (i32, [i64, C.POINTER(vp), vp]),
    "gda_sym_open": (i32, [vp, C.POINTER(vp)]),
    "gda_sym_close": (i32, [vp]),
    "gda_sym_free": (i32, [vp]),
    "gda_peer_barrier": (i32, [vp, i32, i32, u64, vp, vp]),
    "gda_gemm_workspace_bytes": (i64, [i32, i32, i64, i64, i64]),
    "gda_gemm_f32": (i32, [i32, i32, i64, i64, i64, f32, vp, i64, vp, i64, f32, vp, i64, vp, i64, vp]),
    "gda_split_bf16": (i32, [vp, i64, i64, i64, vp, vp, i64, vp]),
    "gda_gemm_bf16x3_supported": (i32, [i64, i64, i64, i64, i64]),
    "gda_gemm_bf16x3_workspace_bytes": (i64, [i64, i64, i64]),
    "gda_gemm_bf16x3": (i32, [i32, i32, i64, i64, i64, vp, vp, i64, vp, vp, i64, vp, i64, vp, i64, vp]),
    "gda_xt_ptr_entries": (i64, [i64, i64]),
    "gda_gemm_xt_fwd": (i32, [vp, vp, vp, vp, i64, i64, i32, i64, vp, vp, i64, vp, i64, vp]),
    "gda_gemm_xt_dw_workspace_bytes": (i64, [i64, i64, i64]),
    "gda_gemm_xt_dw": (i32, [vp, vp, vp, vp, i64, i64, i64, vp, vp, i64, i32, vp, i64, vp, i64, vp]),
    "gda_gemm_bf16_workspace_bytes": (i64, [i64, i64, i64, i32]),
    "gda_gemm_bf16": (i32, [i32, i32, i64, i64, i64, vp, i64, vp, i64, vp, i64, i32, vp, i64, vp]),
    "gda_cast_f32_bf16": (i32, [vp, vp, i64, vp]),
    "gda_cast_bf16_f32": (i32, [vp, vp, i64, vp]),
    "gda_act_dropout_bf16_fwd": (i32, [vp, vp, i64, i32, f32, u64, vp, vp]),
    "gda_act_dropout_bf16_bwd": (i32, [vp, vp, vp, i64, i32, f32, u64, vp, vp]),
    "gda_colsum_bf16": (i32, [vp, i64, i64, i64, vp, vp]),
    "gda_bern_axpy_f32": (i32, [vp, vp, i64, f32, f32, vp, vp]),
    "gda_bern_dtemp_f32": (i32, [vp, vp, i64, f32, vp, vp, vp, vp]),
    "gda_bias_act_dropout_fwd": (i32, [vp, vp, vp, i64, i64, i32, f32, u64, vp, vp]),
    "gda_bias_act_dropout_bwd": (i32, [vp, vp, vp, vp, i64, i64, i32, f32, u64, vp, vp]),
    "gda_bias_act_dropout_rep_fwd": (i32, [vp, vp, vp, i64, i64, i32, i32, f32, u64, vp, vp]),
    "gda_bias_act_dropout_rep_bwd": (i32, [vp, vp, vp, vp, vp, i64, i64, i32, i32, i32, f32, u64, vp, vp]),
    "gda_colsum_f32": (i32, [vp, i64, i64, i64, vp, vp]),
    "gda_softmax_ce_fwd_bwd": (i32, [vp, i64, i32, i64, vp, i64, vp, vp, vp]),
    "gda_softmax_entropy_fwd_bwd": (i32, [vp, i64, i32, i64, vp, vp, vp]),
    "gda_mmd_workspace_bytes": (i64, [i32, i32, i32]),
    "gda_mmd_fwd": (i32, [vp, i64, vp, i64, i32, vp, vp, i32, i32, f32, i32, vp, vp, i64, vp]),
    "gda_mmd_bwd": (i32, [vp, i64, vp, i64, i32, vp, vp, i32, i32, vp, vp, i64, vp, i64, vp, i64, vp]),
    "gda_gat_fwd": (i32, [vp, vp, i32, vp, vp, f32, vp, vp, vp]),
    "gda_gat_bwd": (i32, [vp, vp, i32, vp, vp, f32, vp, vp, vp, vp, vp, vp, vp]),
    "gda_khop_create": (i32, [vp, i64, i64, i32, vp, C.POINTER(vp)]),
    "gda_rw_create": (i32, [vp, i64, i64, i32, u64, vp, C.POINTER(vp)]),
    "gda_edges_size": (i64, [vp]),
    "gda_edges_export": (i32, [vp, vp, vp]),
    "gda_edges_destroy": (i32, [vp]),
    "gda_graph_degrees": (i32, [vp, vp, vp, vp]),
    "gda_row_scale_rsqrt_f32": (i32, [vp, i64, vp, vp, i64, i32, vp]),
    "gda_laplacian_workspace_bytes": (i64, []),
    "gda_laplacian_finish_f32": (i32, [vp, vp, vp, vp, vp, i64, i32, vp, vp, vp, i64, vp]),
    "gda_attention2_fwd": (i32, [vp, vp, i64, i32, vp, vp, vp, vp, vp]),
    "gda_attention2_bwd": (i32, [vp, vp, i64, i32, vp, vp, vp, vp, vp, vp, vp]),
    "gda_ppmi_create": (i32, [vp, i64, i64, i32, i32, u64, vp, C.POINTER(vp)]),
    "gda_ppmi_walks": (i32, [vp, i64, i64, i32, i32, u64, vp, vp]),
    "gda_wedges_size": (i64, [vp]),
    "gda_wedges_export": (i32, [vp, vp, vp, vp, vp]),
    "gda_wedges_destroy": (i32, [vp]),
    "gda_collate_graphs": (i32, [vp, i32, vp, i64, vp, vp, vp, i64, vp, vp, i64, i64, vp, vp, vp, vp]),
    "gda_unpack_rows_delta_f32": (i32, [vp, vp, vp, vp, i64, i64, vp, i64, vp, vp]),
    "gda_unpack_tiles_f32": (i32, [vp, vp, vp, vp, i64, i64, vp, i64, vp, vp]),
    "gda_unpack_values_f32": (i32, [vp, vp, vp, vp, vp, i64, i64, vp, vp]),
    "gda_unpack_rows_f32": (i32, [vp, vp, i32, vp, i64, i64, vp, i64, vp]),
    "gda_argmax_confusion": (i32, [vp, i64, i32, i64, vp, vp, vp, vp, vp]),
    "gda_segment_mean_fwd": (i32, [vp, i64, vp, i64, i32, vp, vp]),
    "gda_segment_mean_bwd": (i32, [vp, vp, i64, i32, vp, i64, vp]),
    "gda_adam_step": (i32, [i32, vp, vp, vp, vp, vp, f32, f32, f32, f32, f32, vp, vp]),
    "gda_fill_f32": (i32, [vp, i64, f32, vp]),
    "gda_axpy_f32": (i32, [vp, vp, i64, f32, vp]),
    "gda_scale_f32": (i32, [vp, vp, i64, f32, vp]),
    "gda_scale_dev_f32": (i32, [vp, vp, i64, f32, vp, vp]),
    "gda_counter_inc": (i32, [vp, vp]),
    "gda_combine_scalars": (i32, [i32, vp, vp, vp, vp]),
}

# calls whose int return is a status code to check
_STATUS = {k for k, (r, _) in SIGNATURES.items() if r is i32 and k not in ("gda_version", "gda_sm_arch", "gda_gemm_bf16x3_supported", "gda_graph_unit_weights")}


def load():
    """Load libgda.so once; raises GdaError with a build hint if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GdaError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C pygda_b200/csrc` (sm_100a, nvcc 12.9). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the header and the build drift apart
        fn.restype = res
        fn.argtypes = args
    if lib.gda_sm_arch() != 100:
        raise GdaError("libgda.so was not built for sm_100a")
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().gda_last_error().decode("utf-8", "replace")
        raise GdaError(f"{what} failed (code {rc}): {msg}")


NVTX = os.environ.get("GDA_NVTX") == "1"      # one NVTX range per C-ABI call (timelines: nsys / ncu --nvtx)


class _Caller:
    """``gda.<name>(...)`` with status checking: ``from pygda_b200._lib import gda``."""

    def __getattr__(self, name):
        full = "gda_" + name
        fn = getattr(load(), full)
        if full in _STATUS:
            if NVTX:
                import torch

                def wrapped(*a, _fn=fn, _n=full, _push=torch.cuda.nvtx.range_push, _pop=torch.cuda.nvtx.range_pop):
                    _push(_n)
                    try:
                        check(_fn(*a), _n)
                    finally:
                        _pop()
            else:
                def wrapped(*a, _fn=fn, _n=full):
                    check(_fn(*a), _n)
            setattr(self, name, wrapped)
            return wrapped
        setattr(self, name, fn)
        return fn


gda = _Caller()
