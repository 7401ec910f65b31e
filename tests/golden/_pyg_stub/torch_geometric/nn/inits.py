from oracle.pyg_ops import glorot_ as _g


def glorot(t):
    if t is not None:
        _g(t)


def zeros(t):
    if t is not None:
        t.data.fill_(0.)


def uniform(size, t):
    import math
    if t is not None:
        bound = 1.0 / math.sqrt(size)
        t.data.uniform_(-bound, bound)
