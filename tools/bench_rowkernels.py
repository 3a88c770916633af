"""A/B timing of the row-streaming kernels (LayerNorm fwd/bwd, activation fwd/bwd) under the PR_TUNE variants, and of the
cuBLAS forms of the weight-gradient GEMM.  CUDA events, L2 flushed between iterations.
    PR_TUNE=1 python tools/bench_rowkernels.py --json gpurun_out/rowk_tune1.json [--gemm]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pixelrec_b200 import ops  # noqa: E402


ITERS, WARM = 12, 3


def timeit(fn, flush):
    iters, warm = ITERS, WARM
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.add_(1.0)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return float(np.median(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=4096)
    ap.add_argument("--L", type=int, default=20)
    ap.add_argument("--D", type=int, default=512)
    ap.add_argument("--json", default=None)
    ap.add_argument("--gemm", action="store_true")
    ap.add_argument("--tunes", default="0,1,2,3")
    ap.add_argument("--once", action="store_true", help="one launch per case, no warm-up (for ncu captures)")
    a = ap.parse_args()
    global ITERS, WARM
    if a.once:
        ITERS, WARM = 1, 0
    dev = torch.device("cuda", 0)
    torch.backends.cuda.matmul.allow_tf32 = True
    B, L, D = a.B, a.L, a.D
    M = B * L
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    out = {"B": B, "L": L, "D": D}
    from pixelrec_b200 import lib as _lib
    L_ = _lib.load()

    def rec(name, ms, nbytes=None, flops=None):
        r = {"ms": ms}
        if nbytes:
            r["GBps"] = nbytes / ms / 1e6
        if flops:
            r["TFLOPs"] = flops / ms / 1e9
        out[name] = r
        print(name, json.dumps(r), flush=True)

    h = torch.randn(B, L, D, device=dev)
    res = torch.randn(B, L, D, device=dev)
    dy = torch.randn(B, L, D, device=dev)
    gam = torch.rand(D, device=dev) + 0.5
    bet = torch.randn(D, device=dev)
    for t in [int(v) for v in a.tunes.split(",")]:
        L_.pr_set_tuning(t)
        y, mean, rstd = ops._raw_add_ln_fwd(h, res, gam, bet, 1e-12, 0.1, 7, 3)
        rec(f"tune{t}_add_ln_fwd_p0.1", timeit(lambda: ops._raw_add_ln_fwd(h, res, gam, bet, 1e-12, 0.1, 7, 3), flush), 3 * M * D * 4)
        rec(f"tune{t}_add_ln_bwd_bias_p0.1", timeit(lambda: ops._raw_add_ln_bwd_bias(dy, h, res, gam, mean, rstd, 0.1, 7, 3), flush), 5 * M * D * 4)
        rec(f"tune{t}_add_ln_bwd_bias_p0", timeit(lambda: ops._raw_add_ln_bwd_bias(dy, h, res, gam, mean, rstd, 0.0, 7, 3), flush), 4 * M * D * 4)
    L_.pr_set_tuning(int(os.environ.get("PR_TUNE", "0")))
    x1 = torch.randn(M, 2 * D, device=dev)
    d1 = torch.randn(M, 2 * D, device=dev)
    rec("act_fwd_gelu", timeit(lambda: ops.activation(x1, "gelu"), flush), 2 * M * 2 * D * 4)
    rec("act_bwd_bias_gelu", timeit(lambda: ops._raw_act_bwd_bias(x1, d1, 0), flush), 3 * M * 2 * D * 4)

    if a.gemm:
        for (o, i) in [(3 * D, D), (D, D), (2 * D, D), (D, 2 * D)]:
            dyo = torch.randn(M, o, device=dev)
            xi = torch.randn(M, i, device=dev)
            fl = 2.0 * M * o * i
            rec(f"wgrad_{o}x{i}_dyT_mm_x", timeit(lambda: dyo.t().mm(xi), flush), flops=fl)
            rec(f"wgrad_{o}x{i}_xT_mm_dy", timeit(lambda: xi.t().mm(dyo), flush), flops=fl)
            for S in (4, 8, 16, 32):
                rec(f"wgrad_{o}x{i}_bmm_split{S}",
                    timeit(lambda: torch.bmm(dyo.view(S, M // S, o).transpose(1, 2), xi.view(S, M // S, i)).sum(0), flush), flops=fl)
            w = torch.randn(o, i, device=dev)
            rec(f"fwd_{o}x{i}", timeit(lambda: xi.mm(w.t()), flush), flops=fl)
            rec(f"dgrad_{o}x{i}", timeit(lambda: dyo.mm(w), flush), flops=fl)
    if a.json:
        with open(a.json, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
