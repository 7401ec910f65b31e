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


def MMD(source_feat, target_feat, sampling_num=1000, times=5, indices=None, kernel_mul=2.0,
        kernel_num=5):
    if indices is None:
        indices = draw_indices(source_feat.size(0), target_feat.size(0), sampling_num, times)
    s_idx, t_idx = indices
    dev = source_feat.device
    s_idx = s_idx.to(dev, non_blocking=True).contiguous()
    t_idx = t_idx.to(dev, non_blocking=True).contiguous()
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
