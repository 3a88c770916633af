"""Per-kernel timing on the GPU box (CUDA events, warm-up, L2 flushed between iterations).
   python tools/microbench.py [--B 4096] [--json gpurun_out/micro.json]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pixelrec_b200 import ops  # noqa: E402


def timeit(fn, iters=10, warm=3, flush=None):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.add_(1.0)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return float(np.median(ts)), float(np.min(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=4096)
    ap.add_argument("--N", type=int, default=97001)
    ap.add_argument("--D", type=int, default=512)
    ap.add_argument("--L", type=int, default=20)
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    B, N, D, L, h = a.B, a.N, a.D, a.L, 4
    g = np.random.default_rng(0)
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)      # 256 MB > 126 MB L2
    res = {}

    def rec(name, ms, bytes_=None, flops=None):
        med, mn = ms
        r = {"ms_median": med, "ms_min": mn}
        if bytes_:
            r["GBps"] = bytes_ / med / 1e6
        if flops:
            r["TFLOPs"] = flops / med / 1e9
        res[name] = r
        print(name, json.dumps(r), flush=True)

    W = torch.randn(N, D, device=dev) * 0.02
    from tests.gpu_util import zipf_ids
    idx_np = zipf_ids(g, B * 2 * (L + 1), N).reshape(B, 2, L + 1)
    idx = torch.from_numpy(idx_np).to(dev)
    R = idx.numel()
    out = None
    for impl in (1, 2):
        rec(f"gather_impl{impl}", timeit(lambda: ops.gather_rows(W, idx, impl=impl), flush=flush), bytes_=R * (8 * D + 8))
    rec("torch_index_select", timeit(lambda: W[idx.view(-1)], flush=flush), bytes_=R * (8 * D + 8))
    src_copy = torch.randn(R, D, device=dev)
    dst_copy = torch.empty(R, D, device=dev)
    rec("torch_copy_same_bytes", timeit(lambda: dst_copy.copy_(src_copy), flush=flush), bytes_=R * 8 * D)

    dE = torch.randn(R, D, device=dev)
    rec("scatter_plan", timeit(lambda: ops.ScatterPlan(idx, N, 0), flush=flush))
    plan = ops.ScatterPlan(idx, N, 0)
    U = plan.n_uniq.item()
    rec("scatter_add", timeit(lambda: ops.scatter_add_rows(dE, plan), flush=flush), bytes_=R * D * 4 + U * D * 4 + R * 8)
    res["scatter_add"]["U"] = U
    M, V = torch.zeros_like(W), torch.zeros_like(W)
    row2slot = torch.full((N,), -1, dtype=torch.int32, device=dev)
    rows = ops.scatter_add_rows(dE, plan)
    rec("adamw_rows_nograd", timeit(lambda: ops.adamw_rows(W, M, V, None, None, 1e-4, 0.9, 0.999, 1e-8, 0.1, 1), flush=flush), bytes_=6 * N * D * 4)
    w2 = torch.randn(8 * D * D * 2, device=dev)
    rec("adamw_dense_4.2M", timeit(lambda: ops.adamw_dense(w2, w2.clone(), torch.zeros_like(w2), torch.zeros_like(w2), 1e-4, 0.9, 0.999, 1e-8, 0.1, 1)))

    x = torch.randn(B, L, D, device=dev, requires_grad=True)
    r_ = torch.randn(B, L, D, device=dev, requires_grad=True)
    gam, bet = torch.ones(D, device=dev, requires_grad=True), torch.zeros(D, device=dev, requires_grad=True)
    rec("add_ln_fwd_p0.1", timeit(lambda: ops.add_ln(x.detach(), r_.detach(), gam.detach(), bet.detach(), 1e-12, p_pre=0.1, seed=1), flush=flush), bytes_=3 * B * L * D * 4)
    y = ops.add_ln(x, r_, gam, bet, 1e-12, p_pre=0.1, seed=1)
    dy = torch.randn_like(y)
    rec("add_ln_bwd_p0.1", timeit(lambda: torch.autograd.grad(y, (x, r_, gam, bet), dy, retain_graph=True), flush=flush), bytes_=5 * B * L * D * 4)
    rec("torch_layernorm_fwd", timeit(lambda: torch.nn.functional.layer_norm(x.detach() + r_.detach(), (D,), gam.detach(), bet.detach(), 1e-12), flush=flush))

    qkv = torch.randn(B, L, 3 * D, device=dev, requires_grad=True)
    ids = torch.ones(B, L, dtype=torch.int64, device=dev)
    for p in (0.0, 0.1):
        rec(f"attn_fwd_p{p}", timeit(lambda: ops.attention(qkv.detach(), ids, h, True, p, 1, 1), flush=flush), bytes_=16 * B * L * D, flops=4 * B * L * L * D)
        c = ops.attention(qkv, ids, h, True, p, 1, 1)
        dc = torch.randn_like(c)
        rec(f"attn_bwd_p{p}", timeit(lambda: torch.autograd.grad(c, qkv, dc, retain_graph=True), flush=flush), bytes_=32 * B * L * D)
    q4 = qkv.detach().view(B, L, 3, h, D // h).permute(2, 0, 3, 1, 4).contiguous()
    rec("torch_sdpa_fwd", timeit(lambda: torch.nn.functional.scaled_dot_product_attention(q4[0], q4[1], q4[2], is_causal=True), flush=flush))

    o = torch.randn(B, L, D, device=dev, requires_grad=True)
    E = torch.randn(B, 2, L + 1, D, device=dev, requires_grad=True)
    rec("bpr_fwd", timeit(lambda: ops.bpr_loss(o.detach(), E.detach(), ids), flush=flush), bytes_=3 * B * L * D * 4)
    ls = ops.bpr_loss(o, E, ids)
    rec("bpr_bwd", timeit(lambda: torch.autograd.grad(ls, (o, E), retain_graph=True), flush=flush), bytes_=6 * B * L * D * 4)

    a1 = torch.randn(B * L, D, device=dev)
    w1 = torch.randn(3 * D, D, device=dev)
    for tf32 in (True, False):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        rec(f"cublas_qkv_gemm_tf32={tf32}", timeit(lambda: a1 @ w1.t()), flops=2 * B * L * D * 3 * D)
    torch.backends.cuda.matmul.allow_tf32 = True
    a16, w16 = a1.bfloat16(), w1.bfloat16()
    rec("cublas_qkv_gemm_bf16", timeit(lambda: a16 @ w16.t()), flops=2 * B * L * D * 3 * D)

    # ---- whole training step (model + fused optimizer)
    from pixelrec_b200.model.IDNet.sasrec import SASRec
    from pixelrec_b200.trainer.optim import FusedAdamW

    class Dl:
        item_num = N
    cfg = dict(n_layers=2, n_heads=4, embedding_size=D, inner_size=2, hidden_dropout_prob=0.1, attn_dropout_prob=0.1,
               hidden_act="gelu", layer_norm_eps=1e-12, initializer_range=0.02, MAX_ITEM_LIST_LENGTH=L, seed=2020)
    m = SASRec(cfg, Dl()).to(dev).train()
    opt = FusedAdamW(m.parameters(), lr=1e-4, weight_decay=0.1, tables=[m.item_embedding])
    mask = (idx[:, 0, 1:] != 0).long()

    def step():
        opt.zero_grad()
        loss = m((idx, mask))
        loss.backward()
        opt.step()
        return loss
    rec("train_step", timeit(step, iters=10, warm=3))
    res["train_step"]["seq_per_s"] = B / res["train_step"]["ms_median"] * 1e3
    print("seq/s", res["train_step"]["seq_per_s"])
    if a.json:
        os.makedirs(os.path.dirname(a.json), exist_ok=True)
        json.dump(res, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
