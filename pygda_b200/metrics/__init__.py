"""``eval_micro_f1`` / ``eval_macro_f1`` with the reference's signatures
(pygda/metrics/metrics.py): labels and predictions go to the host and sklearn scores
them -- host-side by design, outside the accelerated path."""
from sklearn.metrics import f1_score


def eval_micro_f1(label, pred):
    return f1_score(label.cpu().numpy(), pred.cpu().numpy(), average='micro')


def eval_macro_f1(label, pred):
    return f1_score(label.cpu().numpy(), pred.cpu().numpy(), average='macro')


__all__ = ["eval_micro_f1", "eval_macro_f1"]
