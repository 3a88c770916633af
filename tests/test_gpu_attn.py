"""GPU parity: warp-specialised TMA-fed attention core (K4+K6) vs the oracle (layers.py:590-612 +
sasrec.py:119-126), forward and backward, with key padding, causal mask and Philox dropout."""
import numpy as np
import pytest
import torch

from oracle import philox_np as PH
from oracle import sasrec_np as O
from tests.gpu_util import dev, rel, t

pytestmark = pytest.mark.gpu
TOL = 3e-5


def _case(B, L, h, dh, p, seed, pad=True, causal=True):
    g = np.random.default_rng(B * 1000 + L * 10 + dh)
    D = h * dh
    qkv = g.standard_normal((B, L, 3 * D)).astype(np.float32)
    ids = np.ones((B, L), dtype=np.int64)
    if pad:
        for b in range(B):
            n = int(g.integers(1, L + 1))
            ids[b, :L - n] = 0                                 # left padding, >= 1 valid key
        ids[0] = 1
    dctx = g.standard_normal((B, L, D)).astype(np.float32)
    # Fully-masked (left-pad) query rows are don't-care: in fp32 the reference's -1e9 swallows the scores
    # (uniform row), in the float64 oracle it does not.  They never receive gradient in the model (their
    # loss positions are masked, sasrec.py:91), so the test feeds zero upstream gradient there.
    dctx[ids == 0] = 0
    mask = O.attention_mask(ids, np.float64) if causal else np.where((ids != 0)[:, None, None, :], 0.0, -1e9) * np.ones((1, 1, L, 1))
    drop = PH.attn_keep_scale(B, h, L, p, seed, 5).astype(np.float64)
    q, k, v = (qkv[..., i * D:(i + 1) * D].astype(np.float64) for i in range(3))
    ctx_ref, cache = O.attn_core_fwd(q, k, v, mask, h, drop if p > 0 else None)
    dq, dk, dv = O.attn_core_bwd(dctx.astype(np.float64), cache)
    return qkv, ids, dctx, ctx_ref, cache[3], np.concatenate([dq, dk, dv], -1)


@pytest.mark.parametrize("B,L,h,dh", [(3, 10, 4, 32), (5, 20, 4, 128), (2, 7, 2, 128), (4, 20, 4, 16), (2, 12, 4, 64),
                                      (3, 20, 4, 512), (2, 10, 2, 256), (3, 32, 2, 64), (70, 20, 4, 128), (2, 1, 4, 32),
                                      (2, 50, 12, 64)])
@pytest.mark.parametrize("p", [0.0, 0.1])
def test_attention_fwd_bwd(B, L, h, dh, p):
    from pixelrec_b200 import ops
    seed = 99
    qkv, ids, dctx, ctx_ref, p_ref, dqkv_ref = _case(B, L, h, dh, p, seed)
    tq = t(qkv).requires_grad_()
    ctx = ops.attention(tq, t(ids), h, True, p, seed, 5, tf32=False)
    ctx.backward(t(dctx))
    torch.cuda.synchronize()
    valid = ids.astype(bool)                                   # fully-masked query rows are don't-care
    got = ctx.detach().cpu().numpy()
    assert np.isfinite(got).all()
    assert rel(got[valid], ctx_ref[valid]) < TOL
    assert rel(tq.grad.cpu().numpy(), dqkv_ref) < TOL * 3


@pytest.mark.parametrize("B,L,h,dh", [(3, 10, 4, 32), (4, 20, 4, 128)])
def test_attention_masked_rows_match_fp32_reference_arithmetic(B, L, h, dh):
    """With the oracle run in float32 (the reference's precision) the -1e9 arithmetic is identical, so even the
    don't-care rows and the gradient flowing out of them must agree."""
    from pixelrec_b200 import ops
    g = np.random.default_rng(5)
    D = h * dh
    qkv = g.standard_normal((B, L, 3 * D)).astype(np.float32)
    ids = np.ones((B, L), dtype=np.int64)
    ids[0, :L // 2] = 0
    ids[1, :L - 1] = 0
    dctx = g.standard_normal((B, L, D)).astype(np.float32)
    q, k, v = (np.ascontiguousarray(qkv[..., i * D:(i + 1) * D]) for i in range(3))
    ctx_ref, cache = O.attn_core_fwd(q, k, v, O.attention_mask(ids, np.float32), h)
    dq, dk, dv = O.attn_core_bwd(dctx, cache)
    tq = t(qkv).requires_grad_()
    ctx = ops.attention(tq, t(ids), h, True, 0.0, 0, 0, tf32=False)
    ctx.backward(t(dctx))
    assert rel(ctx.detach().cpu().numpy(), ctx_ref) < 1e-5
    assert rel(tq.grad.cpu().numpy(), np.concatenate([dq, dk, dv], -1)) < 1e-4


def test_attention_fully_masked_rows_uniform_never_nan():
    """Left-padded query rows: every key gets -1e9 -> the reference's softmax is uniform over all L keys."""
    from pixelrec_b200 import ops
    B, L, h, dh = 2, 10, 4, 32
    g = np.random.default_rng(0)
    qkv = t(g.standard_normal((B, L, 3 * h * dh)).astype(np.float32))
    ids = np.ones((B, L), dtype=np.int64)
    ids[0, :6] = 0
    ids[1, :] = 0                                              # a completely empty sequence
    out = ops.AttnFn.apply(qkv, t(ids), h, True, 0.0, 0, 0, False)
    probs_ref = O.softmax_lastdim(np.zeros((L,)) - 1e9)
    assert torch.isfinite(out).all()
    v = qkv[..., 2 * h * dh:]
    assert torch.allclose(out[1, 0], v[1].mean(0), atol=1e-5)   # uniform average of all L values
    assert abs(probs_ref.sum() - 1) < 1e-12


def test_attention_bidirectional_no_padding():
    """causal=False, key_ids=None: the ViT item-encoder configuration (PixelNet) of the same kernel."""
    from pixelrec_b200 import ops
    B, L, h, dh = 3, 50, 12, 64
    g = np.random.default_rng(1)
    D = h * dh
    qkv = g.standard_normal((B, L, 3 * D)).astype(np.float32)
    q, k, v = (qkv[..., i * D:(i + 1) * D].astype(np.float64) for i in range(3))
    ref, _ = O.attn_core_fwd(q, k, v, np.zeros((B, 1, L, L)), h)
    out = ops.attention(t(qkv), None, h, False, 0.0, 0, 0, tf32=False)
    assert rel(out.cpu().numpy(), ref) < TOL


def test_attention_rejects_bad_shapes():
    from pixelrec_b200 import ops
    from pixelrec_b200.lib import PixelRecB200Error
    with pytest.raises(PixelRecB200Error):
        ops.attention(torch.zeros(1, 300, 3 * 64, device=dev()), None, 2, True)     # L > 256 (64 < L <= 256: long-sequence kernels)
    with pytest.raises(PixelRecB200Error):
        ops.attention(torch.zeros(1, 8, 3 * 96, device=dev()), None, 2, True)       # dh = 48


# ------------------------------------------------------------------------------------------------ tensor-core path
TOL_TF32 = 2e-3      # TF32 operands (10-bit mantissa), fp32 accumulate; north_star tolerance is 1e-3 on loss/logits


@pytest.fixture(params=[0, 8], ids=["warp_per_item", "warp_pair_per_item"])
def attn_tune(request):
    """tensor-core attention variants (pr_set_tuning bit 8: two warps share one item pipeline); same results required"""
    from pixelrec_b200 import lib
    L_ = lib.load()
    before = L_.pr_set_tuning(-1)
    L_.pr_set_tuning((before & ~8) | request.param)
    yield request.param
    L_.pr_set_tuning(before)


@pytest.mark.parametrize("B,L,h,dh", [(3, 10, 4, 32), (5, 20, 4, 128), (2, 7, 2, 128), (4, 20, 4, 16), (2, 12, 4, 64),
                                      (3, 20, 4, 512), (2, 10, 2, 256), (3, 32, 2, 64), (70, 20, 4, 128), (2, 1, 4, 32),
                                      (300, 20, 4, 128), (4, 16, 4, 96), (3, 24, 2, 128), (33, 17, 4, 32), (9, 32, 4, 256)])
@pytest.mark.parametrize("p", [0.0, 0.1])
def test_attention_tensor_core_path(B, L, h, dh, p, attn_tune):
    """mma.sync TF32 attention core vs the fp64 oracle (same masks, same Philox dropout bits as the fp32 kernel)."""
    from pixelrec_b200 import ops
    if dh % 32:
        pytest.skip("tensor-core path needs dh % 32 == 0 (the fp32 kernel covers the rest)")
    seed = 99
    qkv, ids, dctx, ctx_ref, p_ref, dqkv_ref = _case(B, L, h, dh, p, seed)
    tq = t(qkv).requires_grad_()
    ctx = ops.attention(tq, t(ids), h, True, p, seed, 5, tf32=True)
    ctx.backward(t(dctx))
    torch.cuda.synchronize()
    valid = ids.astype(bool)
    got = ctx.detach().cpu().numpy()
    assert np.isfinite(got).all() and torch.isfinite(tq.grad).all()
    assert rel(got[valid], ctx_ref[valid]) < TOL_TF32
    assert rel(tq.grad.cpu().numpy(), dqkv_ref) < TOL_TF32 * 2


def test_attention_tensor_core_exact_on_tf32_representable_inputs(attn_tune):
    """With operands exactly representable in TF32 the tensor-core path must agree with the fp32 FFMA kernel to fp32
    rounding (1e-6), which pins the fragment layouts independently of TF32 rounding noise."""
    from pixelrec_b200 import ops
    g = np.random.default_rng(4)
    B, L, h, dh = 6, 20, 4, 128
    qkv = (g.integers(-8, 9, size=(B, L, 3 * h * dh)) / 8.0).astype(np.float32)
    ids = np.ones((B, L), dtype=np.int64)
    ids[1, :7] = 0
    a = ops.attention(t(qkv), t(ids), h, True, 0.0, 0, 0, tf32=False)
    b = ops.attention(t(qkv), t(ids), h, True, 0.0, 0, 0, tf32=True)
    valid = torch.from_numpy(ids.astype(bool)).to(a.device)
    # S = QK^T is exact; P is fp32 -> rounded to TF32 before P V, so O carries ~2^-11 relative error
    assert (a - b)[valid].abs().max().item() < 2e-3 * a.abs().max().item()


@pytest.mark.parametrize("B,L,h,dh,p", [(300, 20, 4, 128, 0.1), (40, 32, 2, 64, 0.0), (25, 20, 4, 512, 0.1)])
def test_attention_warp_pair_variant_is_bit_identical(B, L, h, dh, p):
    """Splitting an item between two warps changes who computes a query tile, not the arithmetic: ctx, probs-based grads equal."""
    from pixelrec_b200 import lib, ops
    L_ = lib.load()
    before = L_.pr_set_tuning(-1)
    qkv, ids, dctx, *_ = _case(B, L, h, dh, p, 5)
    outs = []
    try:
        for mask in (before & ~8, before | 8):
            L_.pr_set_tuning(mask)
            tq = t(qkv).requires_grad_()
            ctx = ops.attention(tq, t(ids), h, True, p, 5, 5, tf32=True)
            ctx.backward(t(dctx))
            torch.cuda.synchronize()
            outs.append((ctx.detach().clone(), tq.grad.clone()))
    finally:
        L_.pr_set_tuning(before)
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
