"""Golden output lines of the reference's own ``logger`` (pygda/utils/utility.py:3-115):

    python tests/golden/make_golden_logger.py        # build container only (needs /root/reference)

Writes tests/golden/logger.json: [{"kwargs": {...}, "stdout": "..."}]."""
import contextlib
import io
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_loader import load_reference  # noqa: E402

CASES = [
    dict(epoch=0, loss=1.23456, source_train_acc=0.5, time=0.1234, verbose=2, train=True),
    dict(epoch=17, loss=0.000049, source_train_acc=0.98765, time=12.0, verbose=2, train=True),
    dict(epoch=3, loss=2.5, source_train_acc=None, time=1.5, verbose=2, train=True),
    dict(epoch=3, loss=2.5, source_train_acc=0.25, time=1.5, verbose=1, train=True),
    dict(epoch=3, loss=2.5, source_train_acc=0.25, time=1.5, verbose=0, train=True),
    dict(epoch=1234, loss=(0.5, 1.5), source_train_acc=0.75, time=3.14159, verbose=2, train=True),
    dict(epoch=0, loss=0.75, target=0.625, time=None, verbose=2, train=False),
    dict(epoch=9, loss=0.75, source_train_acc=0.5, target=0.625, time=2.0, verbose=2, train=True),
]


def main():
    ref = load_reference()
    out = []
    for kw in CASES:
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            ref.utility.logger(**kw)
        out.append({"kwargs": {k: (list(v) if isinstance(v, tuple) else v) for k, v in kw.items()}, "stdout": buf.getvalue()})
    json.dump(out, open(os.path.join(HERE, "logger.json"), "w"), indent=1)
    print("wrote logger.json", len(out), "cases")


if __name__ == "__main__":
    main()
