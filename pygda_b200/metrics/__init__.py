"""``eval_micro_f1`` / ``eval_macro_f1`` with the reference's signatures
(pygda/metrics/metrics.py:  labels and predictions go to the host and sklearn scores them), plus the
device-side forms the fit loops use for the per-epoch training score (pygda/models/a2gnn.py:328-329):
``confusion_from_logits`` fuses ``logits.argmax(dim=1)`` with a C x C confusion count on the GPU
(``gda_argmax_confusion``), so one C*C int64 read-back replaces the two N-element D2H copies + sklearn;
``micro_f1_from_logits`` / ``macro_f1_from_logits`` give sklearn's ``f1_score(average=...)`` values from it."""
import ctypes as C

import torch
from sklearn.metrics import average_precision_score, f1_score, roc_auc_score


def eval_micro_f1(label, pred):
    return f1_score(label.cpu().numpy(), pred.cpu().numpy(), average='micro')


def eval_macro_f1(label, pred):
    return f1_score(label.cpu().numpy(), pred.cpu().numpy(), average='macro')


# The remaining host-side wrappers of pygda/metrics/metrics.py (binary scores of other estimators; never on the hot
# path): same signatures and values, so that ``pygda.metrics`` is complete for callers that switch packages.
def eval_roc_auc(label, score):
    """ROC-AUC, mirrored into [0.5, 1] like the reference (metrics.py:4-40)."""
    auc = roc_auc_score(y_true=label.cpu().numpy(), y_score=score.cpu().numpy())
    return 1 - auc if auc < 0.5 else auc


def eval_recall_at_k(label, score, k=None):
    """Fraction of the positives among the k best-scored items (metrics.py:43-80); k defaults to #positives."""
    if k is None:
        k = sum(label)
    return sum(label[score.topk(k).indices]) / sum(label)


def eval_precision_at_k(label, score, k=None):
    """Fraction of positives in the k best-scored items (metrics.py:83-118)."""
    if k is None:
        k = sum(label)
    return sum(label[score.topk(k).indices]) / k


def eval_average_precision(label, score):
    return average_precision_score(y_true=label.cpu().numpy(), y_score=score.cpu().numpy())


def confusion_from_logits(label, logits, return_pred=False):
    """int64 [C, C] confusion counts (row = label, column = argmax prediction) on the HOST; one small D2H."""
    from .._lib import gda
    if not logits.is_cuda:
        raise ValueError("pygda_b200 kernels run on the GPU only (no CPU fallback)")
    z = logits.detach()
    z = z if (z.dtype == torch.float32 and z.stride(-1) == 1) else z.float().contiguous()
    y = label.to(z.device).contiguous()
    rows, c = z.shape
    if y.dtype != torch.int64 or y.numel() != rows:
        raise ValueError("labels must be int64 with one entry per row")
    counts = torch.empty(c, c, dtype=torch.int64, device=z.device)
    bad = torch.empty(1, dtype=torch.int32, device=z.device)
    pred = torch.empty(rows, dtype=torch.int64, device=z.device) if return_pred else None
    gda.argmax_confusion(C.c_void_p(z.data_ptr()), rows, c, z.stride(0), C.c_void_p(y.data_ptr()),
                         C.c_void_p(pred.data_ptr()) if return_pred else None, C.c_void_p(counts.data_ptr()),
                         C.c_void_p(bad.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    host = torch.cat([counts.view(-1), bad.to(torch.int64)]).cpu()
    if int(host[-1]):
        raise ValueError("label outside [0, num_classes)")
    cm = host[:-1].view(c, c)
    return (cm, pred) if return_pred else cm


def f1_from_confusion(cm, average='micro'):
    """sklearn's multiclass ``f1_score(average='micro'|'macro')`` from confusion counts: the classes scored are
    those present in the labels or the predictions; a class with no true and no predicted sample is left out."""
    cm = cm.to(torch.float64)
    tp = cm.diag()
    true_n, pred_n = cm.sum(1), cm.sum(0)
    total = float(cm.sum())
    if average == 'micro':
        return float(tp.sum()) / total if total > 0 else 0.0
    present = (true_n + pred_n) > 0
    denom = true_n + pred_n                      # 2 tp + fp + fn
    f1 = torch.where(denom > 0, 2 * tp / denom.clamp(min=1), torch.zeros_like(tp))
    return float(f1[present].mean()) if bool(present.any()) else 0.0


def micro_f1_from_logits(label, logits):
    return f1_from_confusion(confusion_from_logits(label, logits), 'micro')


def macro_f1_from_logits(label, logits):
    return f1_from_confusion(confusion_from_logits(label, logits), 'macro')


__all__ = ["eval_micro_f1", "eval_macro_f1", "eval_roc_auc", "eval_recall_at_k", "eval_precision_at_k",
           "eval_average_precision", "confusion_from_logits", "f1_from_confusion", "micro_f1_from_logits",
           "macro_f1_from_logits"]
