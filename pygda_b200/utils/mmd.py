"""MMD -- drop-in for pygda/utils/mmd.py (``MMD(source_feat, target_feat,
sampling_num=1000, times=5)``, :109-158).

The sample indices are drawn exactly as the reference draws them -- two
``torch.randint`` calls on the CPU global generator, source first (:148-149) -- so
with the same seed the same rows are compared; the kernels (mmd.cu) never
materialise the n x n x d tensor of :44-46.
"""
import torch

from .. import ops


def draw_indices(source_num, target_num, sampling_num=1000, times=5):
    source_sample = torch.randint(source_num, (times, sampling_num))
    target_sample = torch.randint(target_num, (times, sampling_num))
    return source_sample, target_sample


_pinned = {}


def to_device_async(idx, device):
    """CPU index tensor -> device through a reusable pinned staging buffer: a pageable H2D copy
    would block the host until the GPU drains everything queued before it."""
    if idx.is_cuda:
        return idx.contiguous()
    if torch.device(device).type != "cuda":
        return idx.to(device)                          # nothing to stage: the tensor is not going to a GPU
    key = (tuple(idx.shape), idx.dtype, str(device))
    ring = _pinned.get(key)
    if ring is None:
        ring = _pinned[key] = [[torch.empty(idx.shape, dtype=idx.dtype).pin_memory() for _ in range(4)], 0, []]
    bufs, pos, events = ring
    if len(events) == len(bufs):                     # the buffer about to be reused must have been consumed
        events.pop(0).synchronize()
    buf = bufs[pos]
    ring[1] = (pos + 1) % len(bufs)
    buf.copy_(idx)
    out = buf.to(device, non_blocking=True)
    ev = torch.cuda.Event()
    ev.record()
    events.append(ev)
    return out


def stage_into(dst, idx):
    """Copy a CPU index tensor into the pre-allocated device tensor ``dst`` through the pinned ring
    (used by the CUDA-graph step, whose kernels read the indices from a fixed address)."""
    key = (tuple(idx.shape), idx.dtype, str(dst.device))
    ring = _pinned.get(key)
    if ring is None:
        ring = _pinned[key] = [[torch.empty(idx.shape, dtype=idx.dtype).pin_memory() for _ in range(4)], 0, []]
    bufs, pos, events = ring
    if len(events) == len(bufs):
        events.pop(0).synchronize()
    buf = bufs[pos]
    ring[1] = (pos + 1) % len(bufs)
    buf.copy_(idx)
    dst.copy_(buf, non_blocking=True)
    ev = torch.cuda.Event()
    ev.record()
    events.append(ev)
    return dst


def MMD(source_feat, target_feat, sampling_num=1000, times=5, indices=None, kernel_mul=2.0,
        kernel_num=5):
    if indices is None:
        indices = draw_indices(source_feat.size(0), target_feat.size(0), sampling_num, times)
    s_idx, t_idx = indices
    dev = source_feat.device
    s_idx = to_device_async(s_idx, dev)
    t_idx = to_device_async(t_idx, dev)
    return ops.MMDFn.apply(source_feat, target_feat, s_idx, t_idx, kernel_mul, kernel_num)


def get_MMD(source_feat, target_feat, kernel_mul=2.0, kernel_num=5, fix_sigma=None):
    """pygda/utils/mmd.py:57-107 on all rows (no sampling); needs equal row counts."""
    if fix_sigma:
        raise NotImplementedError("fix_sigma is never set on the reference's hot path")
    ns, nt = source_feat.size(0), target_feat.size(0)
    if ns != nt:
        raise ValueError("get_MMD kernel expects the same number of source and target rows")
    dev = source_feat.device
    idx = torch.arange(ns, device=dev).view(1, -1)
    return ops.MMDFn.apply(source_feat, target_feat, idx, idx.clone(), kernel_mul, kernel_num)
