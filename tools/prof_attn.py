"""minimal driver for ncu: a few launches of the tensor-core attention kernels at C2 (B=4096)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pixelrec_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
B, L, h, dh = 4096, 20, 4, 128
qkv = torch.randn(B, L, 3 * h * dh, device=dev, requires_grad=True)
ids = torch.ones(B, L, dtype=torch.int64, device=dev)
for _ in range(2):
    c = ops.attention(qkv, ids, h, True, 0.1, 1, 1, tf32=True)
    c.backward(torch.randn_like(c))
torch.cuda.synchronize()
