"""PPMI graph construction on libgda: the random-walk co-occurrence graph
``PPMIConv.norm`` builds in pure Python (pygda/nn/ppmi_conv.py:98-172).  See pygda_b200/csrc/ppmi.cu."""
import ctypes as C

import torch

from . import ops
from ._lib import gda, load

ROUNDS = 40            # ppmi_conv.py:134


def _device_edges(edge_index):
    if not edge_index.is_cuda:
        raise ValueError("pygda_b200 builds the PPMI graph on the GPU: edge_index must be a CUDA tensor")
    if edge_index.dtype != torch.int64 or edge_index.dim() != 2 or edge_index.size(0) != 2:
        raise ValueError("edge_index must be int64 [2, E]")
    return edge_index.contiguous()


def _seed(seed):
    # drawn from the CPU generator by default, so torch.manual_seed makes the walks reproducible
    return (int(torch.randint(0, 2 ** 62, (1,)).item()) if seed is None else int(seed)) & 0xFFFFFFFFFFFFFFFF


def ppmi_edges(edge_index, num_nodes, path_len=5, rounds=ROUNDS, seed=None, return_counts=False):
    """``(edge_index [2, M], ppmi scores [M])`` -- every (start, visited) pair of the walks, sorted by
    (start, visited), scores >= 0 with zeros kept (ppmi_conv.py:158-172).  ``return_counts`` adds the visit
    counts [M] (int32)."""
    ei = _device_edges(edge_index)
    h = C.c_void_p(0)
    with torch.cuda.device(ei.device):
        gda.ppmi_create(ops._p(ei), ei.size(1), int(num_nodes), int(path_len), int(rounds), _seed(seed), ops._stream(),
                        C.byref(h))
        try:
            m = int(load().gda_wedges_size(h))
            out = torch.empty(2, m, dtype=torch.int64, device=ei.device)
            w = torch.empty(m, dtype=torch.float32, device=ei.device)
            cnt = torch.empty(m, dtype=torch.int32, device=ei.device) if return_counts else None
            gda.wedges_export(h, ops._p(out), ops._p(w), ops._p(cnt), ops._stream())
            torch.cuda.current_stream(ei.device).synchronize()
        finally:
            load().gda_wedges_destroy(h)
    return (out, w, cnt) if return_counts else (out, w)


def ppmi_walks(edge_index, num_nodes, path_len=5, rounds=ROUNDS, seed=None):
    """The walks behind ``ppmi_edges`` for the same seed: int32 [rounds, N, path_len], -1 past a walk's length."""
    ei = _device_edges(edge_index)
    walks = torch.empty(int(rounds), int(num_nodes), int(path_len), dtype=torch.int32, device=ei.device)
    with torch.cuda.device(ei.device):
        gda.ppmi_walks(ops._p(ei), ei.size(1), int(num_nodes), int(path_len), int(rounds), _seed(seed), ops._p(walks),
                       ops._stream())
    return walks
