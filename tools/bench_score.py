"""Eval scoring: fused tcgen05 score+mask+top-k vs the reference's path (cuBLAS matmul -> index_put(-inf) -> torch.topk)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pixelrec_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
B_e, N, D, k = 1024, 97001, 512, 10
g = np.random.default_rng(0)
seq = torch.randn(B_e, D, device=dev)
W = torch.randn(N, D, device=dev) * 0.02
hu = torch.arange(B_e, device=dev).repeat_interleave(20)
hi = torch.randint(1, N, (B_e * 20,), device=dev)
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


def ref_path():
    sc = seq @ W.t()
    sc[:, 0] = -float("inf")
    sc[hu, hi] = -float("inf")
    return torch.topk(sc, k, dim=-1)


out = {}
for tf32 in (True, False):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    out[f"reference_path_tf32={tf32}_ms"] = timeit(ref_path)
out["fused_tcgen05_ms"] = timeit(lambda: ops.score_topk(seq, W, k, hu, hi))
out["fused_TFLOPs"] = 2 * B_e * N * D / out["fused_tcgen05_ms"] / 1e9
# A/B of the kernel variants (pr_set_tuning bits 16 = v2 epilogue, 32 = + cluster multicast): SCORE_TUNES=0,16,48
if os.environ.get("SCORE_TUNES"):
    from pixelrec_b200 import lib
    L_ = lib.load()
    base = L_.pr_set_tuning(-1)
    i0 = ops.score_topk(seq, W, k, hu, hi)[1].clone()
    for tn in [int(x) for x in os.environ["SCORE_TUNES"].split(",")]:
        L_.pr_set_tuning((base & ~48) | tn)
        ms = timeit(lambda: ops.score_topk(seq, W, k, hu, hi))
        same = bool((ops.score_topk(seq, W, k, hu, hi)[1] == i0).all())
        out[f"tune{tn}_ms"] = ms
        out[f"tune{tn}_TFLOPs"] = 2 * B_e * N * D / ms / 1e9
        out[f"tune{tn}_ids_equal_v1"] = same
    L_.pr_set_tuning(base)
# fp16-operand path (staged): SCORE_F16=1 [with PR_TUNE bit 32 for the multicast]
if os.environ.get("SCORE_F16") == "1":
    W16 = ops.score_prepare_f16(W)
    out["prepare_f16_table_ms"] = timeit(lambda: ops.score_prepare_f16(W))
    ms = timeit(lambda: ops.score_topk_f16(seq, W16, k, hu, hi))
    out["f16_ms"] = ms
    out["f16_TFLOPs"] = 2 * B_e * N * D / ms / 1e9
    i16 = ops.score_topk_f16(seq, W16, k, hu, hi)[1]
    out["f16_idx_agreement_vs_tf32"] = float((i16 == ops.score_topk(seq, W, k, hu, hi)[1]).float().mean())
torch.backends.cuda.matmul.allow_tf32 = False
v1, i1 = ref_path()
v2, i2 = ops.score_topk(seq, W, k, hu, hi)
out["idx_agreement_vs_fp32"] = float((i1 == i2).float().mean())
out["max_abs_val_diff"] = float((v1 - v2).abs().max())
print(json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/bench_score.json", "w"), indent=1)
