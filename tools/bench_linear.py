"""Linear layers of the SASRec step (M = B*L = 81920 rows) on our CTA-pair tcgen05 GEMM vs cuBLAS TF32 (torch.addmm / mm / bmm), CUDA events,
L2 flushed between iterations.  python tools/bench_linear.py [--M 81920] [--only fwd]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pixelrec_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--M", type=int, default=81920)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--out", default="gpurun_out/bench_linear.json")
args = ap.parse_args()
dev = torch.device("cuda", 0)
torch.backends.cuda.matmul.allow_tf32 = True
flush = torch.empty(64 * 1024 * 1024, device=dev)


def timeit(fn, iters=args.iters):
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


M = args.M
out = {"M": M}
# ---- the CTA-pair GEMM (pr_gemm_tf32): forward, input-gradient and weight-gradient forms of every layer shape
g2 = {}
for name, K, N in [("qkv", 512, 1536), ("dense", 512, 512), ("dense_1", 512, 1024), ("dense_2", 1024, 512)]:
    x = torch.randn(M, K, device=dev)
    W = torch.randn(N, K, device=dev) * 0.02
    b = torch.randn(N, device=dev) * 0.05
    dy = torch.randn(M, N, device=dev) * 0.01
    fl = 2.0 * M * N * K
    r = {}
    for tag, ours, ref in [
        ("fwd", lambda: ops.gemm(x, W, bias=b), lambda: torch.addmm(b, x, W.t())),
        ("dgrad", lambda: ops.gemm(dy, W, b_mn=True), lambda: dy.mm(W)),
        ("wgrad", lambda: ops.gemm(dy, x, a_mn=True, b_mn=True, splits=max(1, 72 // (((N + 255) // 256) * ((K + 255) // 256)))),
         lambda: torch.bmm(dy.view(8, M // 8, -1).transpose(1, 2), x.view(8, M // 8, -1)).sum(0)),
    ]:
        try:
            t_o, t_r = timeit(ours), timeit(ref)
            err = float((ours() - ref()).abs().max() / ref().abs().max())
            r[tag] = dict(ours_ms=t_o, ours_TFLOPs=fl / t_o / 1e9, cublas_ms=t_r, cublas_TFLOPs=fl / t_r / 1e9, rel_diff=err)
        except Exception as e:  # noqa: BLE001
            r[tag] = dict(error=str(e)[:200])
    if name == "dense_1":
        try:
            t_o = timeit(lambda: ops.gemm(x, W, bias=b, epi=ops.GEMM_ACT, act="gelu", want_pre=True))
            r["fwd+gelu(pre,out)"] = dict(ours_ms=t_o, ours_TFLOPs=fl / t_o / 1e9)
            h1 = torch.randn(M, N, device=dev)
            dz = torch.randn(M, K, device=dev) * 0.01
            # input gradient of dense_2 (K_contraction = 512 outputs of dense_2 -> N = 1024) through the activation + bias-grad sums
            W2 = torch.randn(K, N, device=dev) * 0.02
            t_o = timeit(lambda: ops.gemm(dz, W2, b_mn=True, aux=h1, epi=ops.GEMM_ACT_BWD, act="gelu", want_colsum=True))
            r["dgrad_dense_2+gelu_bwd+colsum"] = dict(ours_ms=t_o, ours_TFLOPs=fl / t_o / 1e9)
        except Exception as e:  # noqa: BLE001
            r["fused"] = dict(error=str(e)[:200])
    g2[name] = r
    print("gemm", name, r, flush=True)
out["gemm"] = g2
os.makedirs(os.path.dirname(args.out), exist_ok=True)
json.dump(out, open(args.out, "w"), indent=1)
print(json.dumps(out))
