"""Diagnostic dump for pr_gemm_tf32 on integer operands (exact in TF32): for every operand-layout form, which 32x32 output
blocks are wrong and how -- meant for reading layout / descriptor mistakes off a GPU log.  python tools/diag_gemm.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pixelrec_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
g = np.random.default_rng(0)


def run(form, M, N, K, splits=1):
    x = g.integers(-3, 4, size=(M, K)).astype(np.float32)
    W = g.integers(-3, 4, size=(N, K)).astype(np.float32)
    ref = x.astype(np.float64) @ W.astype(np.float64).T
    a_mn, b_mn = form[0] == "t", form[1] == "n" or form == "tt"
    A = torch.from_numpy(np.ascontiguousarray(x.T) if a_mn else x).to(dev)
    B = torch.from_numpy(np.ascontiguousarray(W.T) if b_mn else W).to(dev)
    try:
        got = ops.gemm(A, B, a_mn=a_mn, b_mn=b_mn, splits=splits)
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        print(f"{form} M={M} N={N} K={K} splits={splits}: EXCEPTION {e}", flush=True)
        return
    got = got.cpu().numpy().astype(np.float64)
    bad = got != ref
    print(f"{form} M={M} N={N} K={K} splits={splits}: wrong {bad.mean():.4f}  max|err| {np.abs(got - ref).max():.1f} "
          f"nan {np.isnan(got).mean():.3f}", flush=True)
    if bad.any():
        rb, cb = (M + 31) // 32, (N + 31) // 32
        grid = np.zeros((rb, cb))
        for i in range(rb):
            for j in range(cb):
                grid[i, j] = bad[i * 32:(i + 1) * 32, j * 32:(j + 1) * 32].mean()
        np.set_printoptions(linewidth=250, precision=1, suppress=True)
        print(" wrong fraction per 32x32 block (rows = m blocks, first 16 x 16):\n", grid[:16, :16], flush=True)
        i, j = np.argwhere(bad)[0]
        print(f" first wrong element ({i},{j}): got {got[i, j]} ref {ref[i, j]}; row {i} got[:8] {got[i, :8]} ref[:8] {ref[i, :8]}")
        # does a wrong row equal some other reference row / column permutation?
        for ii in range(min(M, 512)):
            if np.array_equal(got[i], ref[ii]):
                print(f" got row {i} == ref row {ii}")
                break


for form in ("nt", "nn", "tn", "tt"):
    run(form, 256, 256, 32)
    run(form, 256, 256, 128)
    run(form, 512, 512, 256)
run("tt", 512, 512, 512, splits=4)
run("nt", 300, 100, 96)
