"""GPU parity: long-sequence attention core (pr_attn_long_*_f32, 64 < L <= 256: the ViT-B/16 item encoder's 197 tokens) vs the
fp64 oracle (oracle/sasrec_np.py attn_core_fwd/bwd, REC/model/layers.py:590-612).  Tolerance 3e-5 relative-to-max (strict
fp32).  The kernels' logic is also pinned on CPU (tests/test_emu_kernels.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import sasrec_np as O
from tests.gpu_util import rel, t

pytestmark = [pytest.mark.gpu,
]


@pytest.fixture(params=[0, 64], ids=["fp32_fwd", "tf32_mma_fwd"])
def long_variant(request):
    """pr_set_tuning bit 64: forward on mma.sync TF32 (csrc/attn_long_tc.cuh); the backward kernels are the same"""
    from pixelrec_b200 import lib
    L_ = lib.load()
    before = L_.pr_set_tuning(-1)
    L_.pr_set_tuning((before & ~64) | request.param)
    yield request.param
    L_.pr_set_tuning(before)


def _mask(key_ids, causal, B, L):
    valid = np.ones((B, L), bool) if key_ids is None else (key_ids != 0)
    tri = np.tril(np.ones((L, L), bool)) if causal else np.ones((L, L), bool)
    return np.where(valid[:, None, None, :] & tri[None, None], 0.0, -1e9)


@pytest.mark.parametrize("B,L,h,dh,causal,padded", [(3, 197, 12, 64, False, False), (2, 65, 2, 16, False, False),
                                                  (4, 100, 4, 32, True, True), (2, 256, 1, 64, False, True),
                                                  (1, 130, 2, 128, True, False), (5, 77, 3, 8, False, False),
                                                  (6, 50, 12, 64, False, False), (3, 33, 2, 32, True, True), (2, 64, 4, 64, False, True)])
def test_attn_long_fwd_bwd(B, L, h, dh, causal, padded, long_variant):
    from pixelrec_b200 import ops
    tc = bool(long_variant) and dh in (32, 64, 128)
    ftol, gtol = (3e-3, 5e-3) if tc else (3e-5, 1e-4)          # TF32 operands in the tensor-core forward and backward
    g = np.random.default_rng(L * dh + B)
    D = h * dh
    qkv = g.standard_normal((B, L, 3 * D)).astype(np.float32)
    key_ids = None
    if padded:
        key_ids = g.integers(1, 50, size=(B, L)).astype(np.int64)
        key_ids[:, :L // 5] = 0
    q, k, v = (qkv[..., i * D:(i + 1) * D].astype(np.float64) for i in range(3))
    mask = _mask(key_ids, causal, B, L)
    ref, cache = O.attn_core_fwd(q, k, v, mask, h)
    live = (mask == 0).any(-1)[:, 0]                      # rows with no visible key are don't-care
    x = t(qkv).requires_grad_(True)
    out = ops.attention(x, None if key_ids is None else t(key_ids), h, causal=causal)
    got = out.detach().cpu().numpy()
    assert np.isfinite(got).all()
    assert np.abs(got - ref)[live].max() / np.abs(ref).max() < ftol
    dout = g.standard_normal((B, L, D)).astype(np.float32) * live[..., None]
    out.backward(t(dout))
    dq, dk, dv = O.attn_core_bwd(dout.astype(np.float64), cache)
    dx = x.grad.cpu().numpy()
    for i, (want, name) in enumerate(((dq, "dq"), (dk, "dk"), (dv, "dv"))):
        assert rel(dx[..., i * D:(i + 1) * D], want, 1e-6) < gtol, name


def test_attn_long_rejects_dropout_and_bad_shapes():
    from pixelrec_b200 import ops
    from pixelrec_b200.lib import PixelRecB200Error
    x = torch.zeros(1, 100, 3 * 64, device="cuda")
    with pytest.raises(PixelRecB200Error):
        ops.attention(x, None, 4, causal=False, p_drop=0.1)
    with pytest.raises(PixelRecB200Error):
        ops.attention(torch.zeros(1, 300, 3 * 64, device="cuda"), None, 4, causal=False)      # L > 256
    with pytest.raises(PixelRecB200Error):
        ops.attention(torch.zeros(1, 100, 3 * 48, device="cuda"), None, 4, causal=False)      # dh = 12
