"""attention core: fp32 FFMA kernel vs tensor-core (mma.sync TF32) kernel, fwd and bwd."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pixelrec_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
flush = torch.empty(64 * 1024 * 1024, device=dev)


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


from pixelrec_b200 import lib as _lib  # noqa: E402

L_ = _lib.load()
TUNES = [int(v) for v in os.environ.get("ATTN_TUNES", "").split(",") if v]
out = {}
if TUNES:      # A/B of the tensor-core variants only (pr_set_tuning), C2 shape + C3 shape
    base = L_.pr_set_tuning(-1)
    for name, (B, L, h, dh) in {"C2_B4096_L20_dh128": (4096, 20, 4, 128), "C3_B1024_L20_dh512": (1024, 20, 4, 512)}.items():
        D = h * dh
        qkv = torch.randn(B, L, 3 * D, device=dev, requires_grad=True)
        ids = torch.ones(B, L, dtype=torch.int64, device=dev)
        r = {}
        for tn in TUNES:
            L_.pr_set_tuning(tn)
            r[f"tune{tn}_fwd_ms"] = timeit(lambda: ops.attention(qkv.detach(), ids, h, True, 0.1, 1, 1, tf32=True))
            c = ops.attention(qkv, ids, h, True, 0.1, 1, 1, tf32=True)
            dc = torch.randn_like(c)
            r[f"tune{tn}_bwd_ms"] = timeit(lambda: torch.autograd.grad(c, qkv, dc, retain_graph=True))
        out[name] = r
        print(name, json.dumps(r), flush=True)
    L_.pr_set_tuning(base)
    json.dump(out, open("gpurun_out/bench_attn_tunes.json", "w"), indent=1)
    sys.exit(0)
for name, (B, L, h, dh) in {"C2_B4096_L20_dh128": (4096, 20, 4, 128), "C3_B1024_L20_dh512": (1024, 20, 4, 512),
                            "C1_B4096_L10_dh32": (4096, 10, 4, 32)}.items():
    D = h * dh
    qkv = torch.randn(B, L, 3 * D, device=dev, requires_grad=True)
    ids = torch.ones(B, L, dtype=torch.int64, device=dev)
    r = {}
    for tf32 in (False, True):
        tag = "tf32_mma" if tf32 else "fp32_ffma"
        r[tag + "_fwd_ms"] = timeit(lambda: ops.attention(qkv.detach(), ids, h, True, 0.1, 1, 1, tf32=tf32))
        c = ops.attention(qkv, ids, h, True, 0.1, 1, 1, tf32=tf32)
        dc = torch.randn_like(c)
        r[tag + "_bwd_ms"] = timeit(lambda: torch.autograd.grad(c, qkv, dc, retain_graph=True))
        r[tag + "_fwd_GBps"] = 16 * B * L * D / r[tag + "_fwd_ms"] / 1e6
        r[tag + "_bwd_GBps"] = 32 * B * L * D / r[tag + "_bwd_ms"] / 1e6
    out[name] = r
    print(name, json.dumps(r), flush=True)
json.dump(out, open("gpurun_out/bench_attn.json", "w"), indent=1)
