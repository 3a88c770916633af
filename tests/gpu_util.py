import numpy as np
import pytest
import torch

requires_gpu = pytest.mark.gpu


def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)


def t(x, d=None):
    return torch.from_numpy(np.ascontiguousarray(x)).to(d or dev())


def rel(a, b, floor=1e-7):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), floor))


def zipf_ids(g, n, N, alpha=0.8, c=10.0):
    """long-tail item ids in [1,N), randomly permuted so hot rows are not contiguous"""
    r = np.arange(1, N, dtype=np.float64)
    p = 1.0 / (r + c) ** alpha
    p /= p.sum()
    perm = g.permutation(N - 1) + 1
    return perm[g.choice(N - 1, size=n, p=p)].astype(np.int64)
