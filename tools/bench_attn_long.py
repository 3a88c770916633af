"""Long-sequence attention (pr_attn_long_*_f32) at the ViT-B/16 item-encoder shape: fp32 forward vs tensor-core forward
(pr_set_tuning bit 64) and the backward; CUDA events, L2 flushed between iterations.
    python tools/bench_attn_long.py [--images 352] [--L 197] [--heads 12] [--dh 64]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pixelrec_b200 import lib, ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--images", type=int, default=352)      # B=16 sequences x 2(L+1) images at MAX_ITEM_LIST_LENGTH=10
ap.add_argument("--L", type=int, default=197)
ap.add_argument("--heads", type=int, default=12)
ap.add_argument("--dh", type=int, default=64)
a = ap.parse_args()
dev = torch.device("cuda", 0)
D = a.heads * a.dh
qkv = torch.randn(a.images, a.L, 3 * D, device=dev, requires_grad=True)
dout = torch.randn(a.images, a.L, D, device=dev)
flush = torch.empty(64 * 1024 * 1024, device=dev)
L_ = lib.load()
base = L_.pr_set_tuning(-1)


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.add_(1.0)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return float(np.median(ts))


out = {"shape": vars(a), "flop_fwd": 4.0 * a.images * a.heads * a.L * a.L * a.dh}
ref = None
for name, tn in (("fp32", 0), ("tf32_mma", 64)):
    L_.pr_set_tuning((base & ~64) | tn)
    with torch.no_grad():
        ms = timeit(lambda: ops.attention(qkv, None, a.heads, causal=False))
        y = ops.attention(qkv, None, a.heads, causal=False)
    out[f"fwd_{name}_ms"] = ms
    out[f"fwd_{name}_TFLOPs"] = out["flop_fwd"] / ms / 1e9
    if ref is None:
        ref = y
    else:
        out["max_abs_diff_vs_fp32"] = float((y - ref).abs().max())
L_.pr_set_tuning(base)
y = ops.attention(qkv, None, a.heads, causal=False)
out["bwd_ms"] = timeit(lambda: torch.autograd.grad(y, qkv, dout, retain_graph=True))
torch.backends.cuda.matmul.allow_tf32 = False
out["torch_sdpa_fwd_ms"] = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(
    *(x.transpose(1, 2) for x in qkv.detach().view(a.images, a.L, 3, a.heads, a.dh).unbind(2))))
print(json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/bench_attn_long.json", "w"), indent=1)
