"""``logger`` with the reference's call signature and line format
(pygda/utils/utility.py:3-115): ``Epoch NNNN: loss x, source acc y, time z``.
The reference's ``verbose > 2`` branch refers to an undefined ``score`` (:96-111) and
cannot run; it is not reproduced."""


def logger(epoch=0, loss=0, source_train_acc=None, source_val_acc=None, target=None, time=None,
           verbose=0, train=True):
    if verbose <= 0:
        return
    msg = "Epoch {:04d}: ".format(epoch) if train else "Test: "
    if isinstance(loss, tuple):
        msg += "Loss I {:.4f} | Loss O {:.4f} | ".format(loss[0], loss[1])
    else:
        msg += "loss {:.4f}, ".format(loss)
    if verbose > 1:
        if source_train_acc is not None:
            msg += "source acc {:.4f}, ".format(source_train_acc)
        if target is not None:
            msg += "target acc {:.4f}, ".format(target)
        if time is not None:
            msg += "time {:.2f}".format(time)
    print(msg)
