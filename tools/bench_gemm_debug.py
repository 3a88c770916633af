"""Timing experiments on pr_gemm_tf32 (PR_GEMM_DEBUG modes; see csrc/gemm.cu): where does a k-block's time go?"""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pixelrec_b200 import ops
dev = torch.device("cuda", 0)
flush = torch.empty(64 * 1024 * 1024, device=dev)
def timeit(fn, iters=8):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(iters):
        flush.add_(1.0)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    return float(np.median(ts))
M = 81920
out = {}
for name, K, N in [("K512_N1536", 512, 1536), ("K1536_N512", 1536, 512), ("K512_N512", 512, 512), ("K4096_N512", 4096, 512)]:
    x = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev) * 0.02
    t = timeit(lambda: ops.gemm(x, W))
    out[name] = dict(ms=t, TFLOPs=2.0 * M * N * K / t / 1e9)
print(json.dumps(dict(debug=os.environ.get("PR_GEMM_DEBUG", "0"), cg=os.environ.get("PR_GEMM_CG", "2"), stages=os.environ.get("PR_GEMM_STAGES", "auto"), **out)))
