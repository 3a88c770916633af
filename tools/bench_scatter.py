"""A/B timing of the table-gradient segment reduce (pr_scatter_add_rows_f32): TMA-staged ring (PR_TUNE bit 256, rows_ring.cuh)
vs the LDG warp-per-run kernel, on bench.py's long-tail batches.  CUDA events, L2 flushed between iterations; the two variants
must agree bit for bit.  Algorithmic bytes = 4D(R_valid + U) + 8R.
    python tools/bench_scatter.py --json gpurun_out/scatter_ab.json"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pixelrec_b200 import lib as _lib  # noqa: E402
from pixelrec_b200 import ops  # noqa: E402


def timeit(fn, flush, iters=15, warm=3):
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
    ap.add_argument("--json", default=None)
    ap.add_argument("--once", action="store_true", help="one launch per case (ncu captures)")
    ap.add_argument("--quick", action="store_true", help="three C2 / C1 cases only, no gather reference point")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    L_ = _lib.load()
    base = L_.pr_set_tuning(-1)
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    cases = [("C2_B4096", 97001, 512, 20, 4096), ("C2_B64", 97001, 512, 20, 64), ("C2_B16384", 97001, 512, 20, 16384),
             ("C3_B1024", 408375, 2048, 20, 1024), ("C5_B1024", 100001, 4096, 10, 1024), ("C1_B4096", 10001, 128, 10, 4096)]
    if a.quick:
        cases = [c for c in cases if c[0] in ("C2_B4096", "C2_B16384", "C1_B4096")]
    out = {}
    for name, N, D, L, B in cases:
        perm, p = bench.popularity(N)
        items, _ = bench.synth_batch(np.random.default_rng(1), B, N, L, perm, p)
        idx = torch.from_numpy(items).to(dev)
        R = idx.numel()
        dE = torch.randn(R, D, device=dev)
        plan = ops.ScatterPlan(idx, N, 0)
        U = int(plan.n_uniq.item())
        valid = int((idx != 0).sum().item())
        alg = 4 * D * (valid + U) + 8 * R
        rec = {"N": N, "D": D, "B": B, "R": R, "valid_rows": valid, "U": U, "alg_bytes": alg}
        res = {}
        for label, mask in (("ring", base | 256), ("ldg", base & ~256)):
            L_.pr_set_tuning(mask)
            rows = ops.scatter_add_rows(dE, plan)
            res[label] = rows[:U].clone()
            ms = None if a.once else timeit(lambda: ops.scatter_add_rows(dE, plan), flush)
            rec[label] = {"ms": ms, "GBps": alg / ms / 1e6 if ms else None}
        rec["bit_identical"] = bool(torch.equal(res["ring"], res["ldg"]))
        rec["plan_ms"] = None if a.once else timeit(lambda: ops.ScatterPlan(idx, N, 0), flush)
        L_.pr_set_tuning(base)
        out[name] = rec
        print(name, json.dumps(rec), flush=True)
    # reference point: what random 2 KB row reads + sequential writes reach on this part (every id distinct, L2 flushed):
    # the gather kernel of the forward on 172 032 distinct rows of a 400 K x 512 table
    if not a.once and not a.quick:
        Wt = torch.randn(400000, 512, device=dev)
        ids = torch.randperm(400000, device=dev)[:172032].contiguous()
        ms = timeit(lambda: ops.gather_rows(Wt, ids), flush)
        out["gather_distinct_rows"] = {"rows": 172032, "D": 512, "ms": ms, "GBps": 2 * 172032 * 2048 / ms / 1e6}
        ids_s = torch.arange(172032, device=dev)
        ms = timeit(lambda: ops.gather_rows(Wt, ids_s), flush)
        out["gather_sequential_rows"] = {"rows": 172032, "D": 512, "ms": ms, "GBps": 2 * 172032 * 2048 / ms / 1e6}
        print(json.dumps({k: out[k] for k in ("gather_distinct_rows", "gather_sequential_rows")}), flush=True)
    if a.json:
        with open(a.json, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
