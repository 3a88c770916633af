"""GPU parity: linear layer on the tcgen05 pipeline (pr_linear_tf32, K5) vs a float64 reference of nn.Linear [+ erf-GELU]
(REC/model/layers.py:586-588, 613, 651-669).  TF32 operands, fp32 accumulation: 2e-3 of max|y| (the tolerance of the cuBLAS TF32
path it would replace).  The pipeline's protocol and indexing are pinned on CPU (tests/test_emu_kernels.py); it has not run on a
GPU yet, so this file is opt-in (PR_EXPERIMENTAL=1)."""
import math
import os

import numpy as np
import pytest
import torch

from tests.gpu_util import t

pytestmark = [pytest.mark.gpu,
]


@pytest.fixture(params=[0, 32], ids=["unicast", "w_multicast"])
def mcast(request):
    from pixelrec_b200 import lib
    L_ = lib.load()
    before = L_.pr_set_tuning(-1)
    L_.pr_set_tuning((before & ~32) | request.param)
    yield request.param
    L_.pr_set_tuning(before)


@pytest.mark.parametrize("M,N,K,act,bias", [(200, 300, 64, "gelu", True), (1024, 1024, 512, "gelu", True), (81920, 1536, 512, None, True),
                                            (4096, 512, 1024, None, True), (130, 256, 32, "relu", False), (7, 8, 32, None, True)])
def test_linear_tc_matches_reference(M, N, K, act, bias, mcast):
    from pixelrec_b200 import ops
    g = np.random.default_rng(M + N + K)
    x = g.standard_normal((M, K)).astype(np.float32)
    W = (0.05 * g.standard_normal((N, K))).astype(np.float32)
    b = g.standard_normal(N).astype(np.float32) if bias else None
    xd, Wd = t(x), t(W)
    ref_pre = (xd.double() @ Wd.double().t() + (t(b).double() if bias else 0.0))
    ref = ref_pre if act is None else (0.5 * ref_pre * (1 + torch.erf(ref_pre / math.sqrt(2.0))) if act == "gelu" else ref_pre.clamp_min(0))
    out, pre = ops.linear_tc(xd, Wd, t(b) if bias else None, act, want_pre=True)
    tol = 2e-3 * float(ref_pre.abs().max())
    assert float((pre.double() - ref_pre).abs().max()) < tol
    assert float((out.double() - ref).abs().max()) < tol


def test_linear_tc_exact_on_tf32_representable_inputs(mcast):
    from pixelrec_b200 import ops
    g = np.random.default_rng(3)
    x = g.integers(-4, 5, size=(513, 96)).astype(np.float32)
    W = g.integers(-4, 5, size=(260, 96)).astype(np.float32)
    b = g.integers(-4, 5, size=260).astype(np.float32)
    out = ops.linear_tc(t(x), t(W), t(b))
    assert np.array_equal(out.cpu().numpy(), x @ W.T + b)
