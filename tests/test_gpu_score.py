"""GPU parity: K9 fused scoring GEMM (tcgen05, TF32) + mask + top-k vs the oracle's
full_sort_topk (trainer.py:334-336 + collector.py:133)."""
import os

import numpy as np
import pytest
import torch

from oracle import sasrec_np as O
from tests.gpu_util import t

pytestmark = pytest.mark.gpu

# Kernel variants of pr_score_topk_f32 (pr_set_tuning bits 16 / 32, csrc/score.cu "v2"): every test runs under all three
# (v2 is the default since round 2; all of them have run on a B200, profiles/r02a_stage1_staged_kernels.md).
_VARIANTS = [0, 16, 48]


@pytest.fixture(params=_VARIANTS, ids=lambda v: {0: "v1", 16: "v2", 48: "v2_mcast"}[v], autouse=True)
def score_variant(request):
    from pixelrec_b200 import lib
    L_ = lib.load()
    before = L_.pr_set_tuning(-1)
    L_.pr_set_tuning((before & ~48) | request.param)
    yield request.param
    L_.pr_set_tuning(before)


def _hist(g, B_e, N, n_per):
    hu = np.repeat(np.arange(B_e), n_per)
    hi = g.integers(1, N, size=B_e * n_per)
    return hu.astype(np.int64), hi.astype(np.int64)


@pytest.mark.parametrize("B_e,N,D,k", [(5, 300, 32, 10), (128, 257, 64, 5), (130, 1000, 128, 10), (300, 5003, 512, 10),
                                       (1024, 20011, 512, 10), (64, 256, 64, 16), (33, 777, 96, 20), (7, 40, 32, 32)])
def test_score_topk_exact_on_tf32_representable_inputs(B_e, N, D, k):
    """Small-integer operands are exact in TF32 and their dot products exact in fp32, so the fused kernel must
    reproduce the oracle bit for bit -- including the tie-breaking (lower item id first) that a stable sort gives.
    Integer scores collide massively, so this exercises ties in the per-thread lists AND in the split merge."""
    from pixelrec_b200 import ops
    g = np.random.default_rng(B_e * 7 + N)
    seq = g.integers(-3, 4, size=(B_e, D)).astype(np.float32)
    W = g.integers(-3, 4, size=(N, D)).astype(np.float32)
    hu, hi = _hist(g, B_e, N, 6)
    scores = seq.astype(np.float64) @ W.astype(np.float64).T
    v_ref, i_ref = O.full_sort_topk(scores, hu, hi, k)
    val, idx = ops.score_topk(t(seq), t(W), k, t(hu), t(hi))
    torch.cuda.synchronize()
    assert np.array_equal(idx.cpu().numpy(), i_ref)
    assert np.array_equal(val.cpu().numpy().astype(np.float64), v_ref)
    assert not (idx == 0).any()


@pytest.mark.parametrize("B_e,N,D", [(64, 3000, 128), (1024, 97001, 512)])
def test_score_topk_gaussian_within_tf32_tolerance(B_e, N, D):
    """Real-valued operands: values within 2e-3 of max|score| (TF32 inputs, fp32 accumulate; north_star asks 1e-3
    relative on fp32 logits -- TF32 is what torch 1.10 ran), ranks may only swap between near-ties."""
    from pixelrec_b200 import ops
    g = np.random.default_rng(N)
    seq = g.standard_normal((B_e, D)).astype(np.float32)
    W = (0.02 * g.standard_normal((N, D))).astype(np.float32)
    hu, hi = _hist(g, B_e, N, 12)
    k = 10
    scores = seq.astype(np.float64) @ W.astype(np.float64).T
    masked = scores.copy()
    masked[:, 0] = -np.inf
    masked[hu, hi] = -np.inf
    v_ref, i_ref = O.full_sort_topk(scores, hu, hi, k)
    val, idx = ops.score_topk(t(seq), t(W), k, t(hu), t(hi))
    val, idx = val.cpu().numpy(), idx.cpu().numpy()
    tol = 2e-3 * np.abs(scores).max()
    assert (np.diff(val, axis=1) <= 0).all()                                  # descending
    assert np.abs(val - np.take_along_axis(masked, idx, 1)).max() < tol          # returned values are the true scores
    assert np.isfinite(np.take_along_axis(masked, idx, 1)).all()                 # never a masked item
    assert (val[:, -1] >= v_ref[:, -1] - tol).all()                              # nothing better was missed
    assert (idx == i_ref).mean() > 0.97                                          # only near-ties may swap


@pytest.mark.parametrize("B_e,N,D", [(64, 3000, 128), (1024, 97001, 512), (300, 5003, 512)])
def test_score_topk_exact_ids_equal_the_fp32_ranking(B_e, N, D):
    """pr_score_topk_exact_f32: the ids (and their order) are those of ranking the fp32 scores -- what the reference's
    torch.topk on an fp32 matmul returns (collector.py:133) -- not of ranking TF32 scores."""
    from pixelrec_b200 import ops
    g = np.random.default_rng(N + 1)
    seq = g.standard_normal((B_e, D)).astype(np.float32)
    W = (0.02 * g.standard_normal((N, D))).astype(np.float32)
    hu, hi = _hist(g, B_e, N, 12)
    k = 10
    scores = seq.astype(np.float64) @ W.astype(np.float64).T
    v_ref, i_ref = O.full_sort_topk(scores, hu, hi, k)
    val, idx, nfb = ops.score_topk_exact(t(seq), t(W), k, t(hu), t(hi))
    val, idx = val.cpu().numpy(), idx.cpu().numpy()
    # fp32 dot products differ from the fp64 oracle by ~1e-6 relative: only pairs closer than that may swap
    diff = idx != i_ref
    if diff.any():
        r, c = np.nonzero(diff)
        gap = np.abs(v_ref[r, c] - scores[r, idx[r, c]])
        assert gap.max() < 1e-5 * np.abs(scores).max(), gap.max()
    assert diff.mean() < 1e-3
    assert np.abs(val - v_ref).max() < 1e-5 * np.abs(scores).max()                 # fp32 values, not TF32 ones
    assert int(nfb.item()) <= B_e // 20                                           # the whole-catalog fallback is the rare path


def test_score_topk_exact_fallback_path_on_crowded_scores():
    """integer operands put hundreds of items on the same score: the candidate list cannot be proven complete, every such row
    goes through the whole-catalog fp32 fallback, and the result must still be the oracle's (ties: lower id first)"""
    from pixelrec_b200 import ops
    g = np.random.default_rng(5)
    B_e, N, D, k = 40, 3000, 64, 10
    W = g.integers(-2, 3, size=(N, D)).astype(np.float32)
    W[500:560] = W[7]                                   # 61 identical items ...
    seq = (2 * W[7][None, :] + g.integers(-1, 2, size=(B_e, D))).astype(np.float32)      # ... that every user scores highest
    hu, hi = _hist(g, B_e, N, 6)
    scores = seq.astype(np.float64) @ W.astype(np.float64).T
    v_ref, i_ref = O.full_sort_topk(scores, hu, hi, k)
    assert (v_ref[:, 0] == v_ref[:, -1]).all()          # the whole top-k is one 61-way tie: 32 candidates cannot settle it
    val, idx, nfb = ops.score_topk_exact(t(seq), t(W), k, t(hu), t(hi))
    assert np.array_equal(idx.cpu().numpy(), i_ref)
    assert np.array_equal(val.cpu().numpy().astype(np.float64), v_ref)
    assert int(nfb.item()) == B_e


def test_score_topk_exact_matches_reference_golden(golden):
    """ids identical to the reference's own masked top-k (goldens: torch.topk on the reference's fp32 predict scores)"""
    from pixelrec_b200 import ops
    from tests.test_gpu_sasrec import build
    torch.backends.cuda.matmul.allow_tf32 = False
    m = build(golden).eval()
    seq = m.encode_last(t(golden["eval_item_seq"]))
    W = m.compute_item_all()
    torch.backends.cuda.matmul.allow_tf32 = True
    if W.shape[1] % 32:
        pytest.skip("D % 32 != 0")
    val, idx, _ = ops.score_topk_exact(seq, W.contiguous(), 10, t(golden["eval_hist_u"]), t(golden["eval_hist_i"]))
    assert np.array_equal(idx.cpu().numpy(), golden["eval_topk_idx"])
    assert np.abs(val.cpu().numpy() - golden["eval_topk_val"]).max() < 1e-5 * np.abs(golden["eval_scores_raw"]).max()


def test_score_topk_matches_reference_golden(golden):
    from pixelrec_b200 import ops
    from tests.test_gpu_sasrec import build
    torch.backends.cuda.matmul.allow_tf32 = False
    m = build(golden).eval()
    seq = m.encode_last(t(golden["eval_item_seq"]))
    W = m.compute_item_all()
    if W.shape[1] % 32:
        pytest.skip("D % 32 != 0")
    val, idx = ops.score_topk(seq, W.contiguous(), 10, t(golden["eval_hist_u"]), t(golden["eval_hist_i"]))
    ref_v, ref_i = golden["eval_topk_val"], golden["eval_topk_idx"]
    tol = 2e-3 * np.abs(golden["eval_scores_raw"]).max()
    assert np.abs(val.cpu().numpy() - ref_v).max() < tol
    agree = (idx.cpu().numpy() == ref_i).mean()
    assert agree > 0.9, agree
    torch.backends.cuda.matmul.allow_tf32 = True


def test_score_topk_fewer_valid_items_than_k_and_bad_args():
    from pixelrec_b200 import ops
    from pixelrec_b200.lib import PixelRecB200Error
    g = np.random.default_rng(0)
    N, D = 12, 32
    seq = t(g.integers(-2, 3, size=(3, D)).astype(np.float32))
    W = t(g.integers(-2, 3, size=(N, D)).astype(np.float32))
    hu = t(np.zeros(9, dtype=np.int64))
    hi = t(np.arange(1, 10, dtype=np.int64))                  # user 0 has only items 10, 11 left
    val, idx = ops.score_topk(seq, W, 5, hu, hi)
    assert set(idx[0, :2].tolist()) == {10, 11} and (idx[0, 2:] == -1).all() and torch.isinf(val[0, 2:]).all()
    assert (idx[1] > 0).all()
    with pytest.raises(PixelRecB200Error):
        ops.score_topk(t(np.zeros((2, 48), np.float32)), t(np.zeros((10, 48), np.float32)), 5)    # D % 32
    with pytest.raises(PixelRecB200Error):
        ops.score_topk(seq, W, 33)


@pytest.mark.parametrize("B_e,N,D", [(64, 3000, 128), (1024, 97001, 512), (5, 300, 32)])
def test_score_ce_matches_oracle(B_e, N, D):
    """full-catalog softmax CE on the v2 scoring pipeline (extension; oracle restates F.cross_entropy): TF32 operands, fp32
    accumulation -> 2e-3 of max|logit| on lse / target logit"""
    from pixelrec_b200 import ops
    g = np.random.default_rng(N + 1)
    seq = g.standard_normal((B_e, D)).astype(np.float32)
    W = (0.05 * g.standard_normal((N, D))).astype(np.float32)
    target = g.integers(1, N, size=B_e).astype(np.int64)
    scores = seq.astype(np.float64) @ W.astype(np.float64).T
    lse_r, tl_r, nll_r = O.full_catalog_ce(scores, target)
    lse, tl, nll = (x.cpu().numpy() for x in ops.score_ce(t(seq), t(W), t(target)))
    tol = 2e-3 * np.abs(scores).max()
    assert np.abs(lse - lse_r).max() < tol and np.abs(tl - tl_r).max() < tol and np.abs(nll - nll_r).max() < 2 * tol


@pytest.mark.parametrize("B_e,N,D,chunk", [(300, 5003, 64, 2048), (128, 1000, 128, 8192), (1024, 20011, 512, 8192), (37, 333, 32, 100)])
def test_score_ce_backward_matches_autograd(B_e, N, D, chunk):
    """ops.score_ce_loss (extension): forward on the fused tcgen05 scoring kernel, backward as chunked recompute on pr_gemm_tf32 +
    pr_ce_grad_chunk_f32 -- gradients of sum_r w_r * nll_r w.r.t. seq_out and the table vs float64 autograd of
    F.cross_entropy-style logits with the padding column excluded.  TF32 operands: 2e-3 of the largest gradient entry."""
    from pixelrec_b200 import ops
    g = np.random.default_rng(N + B_e)
    seq = g.standard_normal((B_e, D)).astype(np.float32)
    W = (0.2 * g.standard_normal((N, D))).astype(np.float32)
    target = g.integers(1, N, size=B_e).astype(np.int64)
    target[0] = N - 1                                               # the last (ragged) chunk holds a target
    wts = g.random(B_e).astype(np.float32) + 0.5
    X64 = torch.from_numpy(seq).double().requires_grad_()
    W64 = torch.from_numpy(W).double().requires_grad_()
    sc = X64 @ W64.t()
    sc = torch.cat([torch.full_like(sc[:, :1], -float("inf")), sc[:, 1:]], 1)
    nll_ref = torch.logsumexp(sc, 1) - sc.gather(1, torch.from_numpy(target)[:, None])[:, 0]
    (nll_ref * torch.from_numpy(wts).double()).sum().backward()
    old = ops.CE_CHUNK
    ops.CE_CHUNK = chunk
    try:
        Xg, Wg = t(seq).requires_grad_(), t(W).requires_grad_()
        nll = ops.score_ce_loss(Xg, Wg, t(target))
        (nll * t(wts)).sum().backward()
    finally:
        ops.CE_CHUNK = old
    assert np.abs(nll.detach().cpu().numpy() - nll_ref.detach().numpy()).max() < 4e-3 * max(1.0, float(sc[:, 1:].abs().max()))
    for got, ref, name in ((Xg.grad, X64.grad, "d seq_out"), (Wg.grad, W64.grad, "d table")):
        err = (got.double().cpu() - ref).abs().max().item()
        assert err < 2e-3 * ref.abs().max().item(), (name, err, ref.abs().max().item())
    assert Wg.grad[0].abs().max().item() < 1e-6                      # the padding item gets no gradient


@pytest.mark.parametrize("B_e,N,D,k", [(5, 300, 64, 10), (300, 5003, 512, 10), (1024, 20011, 512, 10), (33, 777, 128, 20)])
@pytest.mark.parametrize("ares", [0, 128], ids=["ring", "resident_seq"])
def test_score_topk_f16_exact_on_small_integers(B_e, N, D, k, ares):
    """fp16-operand scoring (kind::f16): small integers are exact in fp16 -> bit-exact ids, values and tie order.
    ares (pr_set_tuning bit 128): the seq_out tile stays resident in shared memory, the ring carries table tiles only."""
    from pixelrec_b200 import lib, ops
    L_ = lib.load()
    L_.pr_set_tuning((L_.pr_set_tuning(-1) & ~128) | ares)          # the autouse score_variant fixture restores the mask
    g = np.random.default_rng(B_e * 11 + N)
    seq = g.integers(-3, 4, size=(B_e, D)).astype(np.float32)
    W = g.integers(-3, 4, size=(N, D)).astype(np.float32)
    hu, hi = _hist(g, B_e, N, 6)
    v_ref, i_ref = O.full_sort_topk(seq.astype(np.float64) @ W.astype(np.float64).T, hu, hi, k)
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    W16 = ops.score_prepare_f16(t(W), status)
    assert W16.dtype == torch.float16 and torch.equal(W16.float().cpu(), torch.from_numpy(W))
    val, idx = ops.score_topk_f16(t(seq), W16, k, t(hu), t(hi), status=status)
    assert np.array_equal(idx.cpu().numpy(), i_ref)
    assert np.array_equal(val.cpu().numpy().astype(np.float64), v_ref)
    assert int(status.item()) == 0


def test_score_topk_f16_gaussian_matches_tf32_quality_and_flags_overflow():
    from pixelrec_b200 import ops
    B_e, N, D, k = 1024, 97001, 512, 10
    g = np.random.default_rng(7)
    seq = g.standard_normal((B_e, D)).astype(np.float32)
    W = (0.02 * g.standard_normal((N, D))).astype(np.float32)
    hu, hi = _hist(g, B_e, N, 12)
    scores = seq.astype(np.float64) @ W.astype(np.float64).T
    v_ref, i_ref = O.full_sort_topk(scores, hu, hi, k)
    W16 = ops.score_prepare_f16(t(W))
    assert torch.equal(W16.cpu(), torch.from_numpy(W).half())                     # same rounding as torch's fp16 cast
    val, idx = ops.score_topk_f16(t(seq), W16, k, t(hu), t(hi))
    val, idx = val.cpu().numpy(), idx.cpu().numpy()
    tol = 2e-3 * np.abs(scores).max()
    masked = scores.copy()
    masked[:, 0] = -np.inf
    masked[hu, hi] = -np.inf
    assert np.abs(val - np.take_along_axis(masked, idx, 1)).max() < tol
    assert (val[:, -1] >= v_ref[:, -1] - tol).all() and (idx == i_ref).mean() > 0.97
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    big = t(np.full((4, 64), 1e6, np.float32))
    out = ops.score_prepare_f16(big, status)
    assert int(status.item()) == 4 and torch.isfinite(out).all()
