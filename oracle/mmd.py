"""Oracle restatement of pygda/utils/mmd.py (CPU, plain torch).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Pinned against the reference's
own file by tests/golden (that file imports only torch + numpy, so it runs
here unmodified).
"""
import torch


def pairwise_sqdist_broadcast(total):
    """pygda/utils/mmd.py:44-46: materialises [n, n, d] (reference-faithful)."""
    n, d = total.shape
    t0 = total.unsqueeze(0).expand(n, n, d)
    t1 = total.unsqueeze(1).expand(n, n, d)
    return ((t0 - t1) ** 2).sum(2)


def pairwise_sqdist_blocked(total, block=256):
    """Same values as the broadcast form, computed in row blocks so that the
    temporary is [block, n, d] (identical per-element arithmetic: subtract,
    square, sum over d)."""
    n, _ = total.shape
    out = []
    for s in range(0, n, block):
        blk = total[s:s + block]
        out.append(((total.unsqueeze(0) - blk.unsqueeze(1)) ** 2).sum(2))
    return torch.cat(out, 0)


def gaussian_kernel(source, target, kernel_mul=2.0, kernel_num=5, fix_sigma=None,
                    sqdist=pairwise_sqdist_broadcast):
    """pygda/utils/mmd.py:4-55.  Bandwidth uses ``L2.data`` (stop-gradient, :50),
    is divided by kernel_mul**(kernel_num//2) (:51) and the kernels are summed
    over bandwidth * kernel_mul**i (:52-55)."""
    n = int(source.size(0)) + int(target.size(0))
    total = torch.cat([source, target], dim=0)
    l2 = sqdist(total)
    if fix_sigma:
        bandwidth = fix_sigma
    else:
        bandwidth = (torch.sum(l2.detach()) + 1e-6) / (n ** 2 - n)
    bandwidth = bandwidth / (kernel_mul ** (kernel_num // 2))
    vals = [torch.exp(-l2 / (bandwidth * (kernel_mul ** i))) for i in range(kernel_num)]
    return sum(vals)


def get_mmd(source_feat, target_feat, kernel_mul=2.0, kernel_num=5, fix_sigma=None,
            sqdist=pairwise_sqdist_broadcast):
    """pygda/utils/mmd.py:57-107."""
    k = gaussian_kernel(source_feat, target_feat, kernel_mul, kernel_num, fix_sigma, sqdist)
    b = min(int(source_feat.size(0)), int(target_feat.size(0)))
    return torch.mean(k[:b, :b] + k[b:, b:] - k[:b, b:] - k[b:, :b])


def draw_mmd_indices(source_num, target_num, sampling_num=1000, times=5):
    """pygda/utils/mmd.py:148-149: two ``torch.randint`` draws on the CPU global
    RNG, source first.  (Index part of the path: bit-exact.)"""
    source_sample = torch.randint(source_num, (times, sampling_num))
    target_sample = torch.randint(target_num, (times, sampling_num))
    return source_sample, target_sample


def MMD(source_feat, target_feat, sampling_num=1000, times=5, indices=None,
        sqdist=pairwise_sqdist_broadcast):
    """pygda/utils/mmd.py:109-158.  ``indices`` lets a test inject the sample
    indices; by default they are drawn exactly as the reference draws them."""
    if indices is None:
        indices = draw_mmd_indices(source_feat.size(0), target_feat.size(0), sampling_num, times)
    source_sample, target_sample = indices
    times = source_sample.size(0)
    mmd = 0
    for i in range(times):
        mmd = mmd + get_mmd(source_feat[source_sample[i]], target_feat[target_sample[i]],
                            sqdist=sqdist)
    return mmd / times
