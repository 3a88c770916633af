"""GPU parity: the CTA-pair tcgen05 GEMM (pr_gemm_tf32, csrc/gemm.cu) that carries the encoder's nn.Linear layers
(REC/model/layers.py:586-588, 613, 666, 669) and their two backward GEMMs, against a float64 reference.  TF32 operands with
fp32 accumulation: 2e-3 of max|out| (the tolerance of the cuBLAS TF32 path it replaces; the whole-model tests hold north_star's
1e-3 on the loss and gradients).  Ragged M / N tails, K-major and MN-major operands, every epilogue, split-K."""
import math

import numpy as np
import pytest
import torch

from tests.gpu_util import t

pytestmark = pytest.mark.gpu
TOL = 2e-3


def _rel(got, ref):
    ref = np.asarray(ref, dtype=np.float64)
    return float(np.abs(got.double().cpu().numpy() - ref).max() / max(np.abs(ref).max(), 1e-12))


def _gelu(x):
    return x * 0.5 * (1.0 + np.vectorize(math.erf)(x / math.sqrt(2.0)))


def _dgelu(x):
    cdf = 0.5 * (1.0 + np.vectorize(math.erf)(x / math.sqrt(2.0)))
    return cdf + x * np.exp(-0.5 * x * x) / math.sqrt(2.0 * math.pi)


@pytest.mark.parametrize("M,N,K", [(128, 256, 32), (256, 256, 64), (300, 100, 96), (1000, 1536, 512), (513, 512, 1024),
                                   (4099, 260, 128)])
def test_forward_layout_bias(M, N, K):
    from pixelrec_b200 import ops
    g = np.random.default_rng(M + N + K)
    x, W, b = g.standard_normal((M, K)), g.standard_normal((N, K)) * 0.05, g.standard_normal(N) * 0.1
    ref = x @ W.T + b
    got = ops.gemm(t(x.astype(np.float32)), t(W.astype(np.float32)), bias=t(b.astype(np.float32)))
    assert got.shape == (M, N)
    assert _rel(got, ref) < TOL
    got = ops.gemm(t(x.astype(np.float32)), t(W.astype(np.float32)))
    assert _rel(got, x @ W.T) < TOL


@pytest.mark.parametrize("form", ["nt", "nn", "tn", "tt", "tt_split4"])
@pytest.mark.parametrize("M,N,K", [(640, 512, 256), (256, 256, 32), (132, 36, 64)])
def test_exact_on_tf32_representable_operands(form, M, N, K):
    """small integers are exact in TF32 and their sums in fp32: any indexing / swizzle / pairing mistake shows as a wrong integer.
    nt: forward (both K-major); nn: input-gradient form (B MN-major); tt: weight-gradient form (both MN-major)"""
    from pixelrec_b200 import ops
    g = np.random.default_rng(0)
    x = g.integers(-4, 5, size=(M, K)).astype(np.float32)
    W = g.integers(-4, 5, size=(N, K)).astype(np.float32)
    ref = x.astype(np.float64) @ W.astype(np.float64).T
    a_mn, b_mn = form[0] == "t", form[1] == "n" or form.startswith("tt")
    A = t(np.ascontiguousarray(x.T)) if a_mn else t(x)
    B = t(np.ascontiguousarray(W.T)) if b_mn else t(W)
    splits = 4 if form.endswith("split4") and K >= 128 else 1
    got = ops.gemm(A, B, a_mn=a_mn, b_mn=b_mn, splits=splits).cpu().numpy().astype(np.float64)
    assert np.array_equal(got, ref), f"{(got != ref).mean():.3f} of the outputs differ"


@pytest.mark.parametrize("M,N,K", [(384, 512, 256), (1000, 1024, 512), (130, 36, 64)])
def test_input_gradient_form(M, N, K):
    """dx = dy W: A = dy [M, K=out_features] K-major, B = W [out_features, in_features] read as MN-major [K, N]"""
    from pixelrec_b200 import ops
    g = np.random.default_rng(M * 3 + N)
    dy, W = g.standard_normal((M, K)), g.standard_normal((K, N)) * 0.05
    res = g.standard_normal((M, N))
    got = ops.gemm(t(dy.astype(np.float32)), t(W.astype(np.float32)), b_mn=True)
    assert _rel(got, dy @ W) < TOL
    got = ops.gemm(t(dy.astype(np.float32)), t(W.astype(np.float32)), b_mn=True, aux=t(res.astype(np.float32)), epi=ops.GEMM_ADD)
    assert _rel(got, dy @ W + res) < TOL


@pytest.mark.parametrize("R,O,I,splits", [(512, 256, 256, 1), (4096, 512, 512, 8), (81920, 1536, 512, 6), (3000, 100, 260, 3), (1000, 128, 64, 2),
                                          (2080, 512, 1024, 5)])
def test_weight_gradient_form(R, O, I, splits):
    """dW[O, I] = dy^T x: A = dy [R, O] and B = x [R, I], both MN-major (contraction over the R rows), split-K"""
    from pixelrec_b200 import ops
    g = np.random.default_rng(R + O)
    dy, x = (g.standard_normal((R, O)) * 0.1).astype(np.float32), g.standard_normal((R, I)).astype(np.float32)
    ref = dy.astype(np.float64).T @ x.astype(np.float64)
    got = ops.gemm(t(dy), t(x), a_mn=True, b_mn=True, splits=splits)
    assert got.shape == (O, I)
    assert _rel(got, ref) < TOL
    again = ops.gemm(t(dy), t(x), a_mn=True, b_mn=True, splits=splits)
    assert torch.equal(got, again)                         # fixed-order split-K reduction: deterministic


@pytest.mark.parametrize("M,N,K", [(256, 256, 64), (1000, 1024, 512), (77, 132, 32)])
def test_activation_epilogues(M, N, K):
    from pixelrec_b200 import ops
    g = np.random.default_rng(N + K)
    x, W, b = g.standard_normal((M, K)), g.standard_normal((N, K)) * 0.1, g.standard_normal(N) * 0.1
    pre_ref = x @ W.T + b
    out, pre = ops.gemm(t(x.astype(np.float32)), t(W.astype(np.float32)), bias=t(b.astype(np.float32)), epi=ops.GEMM_ACT,
                        act="gelu", want_pre=True)
    assert _rel(pre, pre_ref) < TOL and _rel(out, _gelu(pre_ref)) < TOL
    out = ops.gemm(t(x.astype(np.float32)), t(W.astype(np.float32)), bias=t(b.astype(np.float32)), epi=ops.GEMM_ACT, act="relu")
    assert _rel(out, np.maximum(pre_ref, 0)) < TOL
    # input gradient through the activation + its column sums (bias gradient of dense_1)
    h1 = g.standard_normal((M, N))
    dy, W2 = g.standard_normal((M, K)), g.standard_normal((K, N)) * 0.1
    ref = (dy @ W2) * _dgelu(h1)
    got, cs = ops.gemm(t(dy.astype(np.float32)), t(W2.astype(np.float32)), b_mn=True, aux=t(h1.astype(np.float32)),
                       epi=ops.GEMM_ACT_BWD, act="gelu", want_colsum=True)
    assert _rel(got, ref) < TOL
    assert _rel(cs, got.double().cpu().numpy().sum(0)) < 1e-5
    assert _rel(cs, ref.sum(0)) < 5 * TOL


def test_full_step_shape_against_cublas():
    """M = 81920 (B=4096, L=20): every tile of the persistent schedule, compared with cuBLAS TF32 on the same operands"""
    from pixelrec_b200 import ops
    torch.manual_seed(0)
    M, K, N = 81920, 512, 1024
    x = torch.randn(M, K, device="cuda")
    W = torch.randn(N, K, device="cuda") * 0.02
    b = torch.randn(N, device="cuda") * 0.05
    torch.backends.cuda.matmul.allow_tf32 = False
    ref = torch.addmm(b, x, W.t())
    torch.backends.cuda.matmul.allow_tf32 = True
    got = ops.gemm(x, W, bias=b)
    assert float((got - ref).abs().max() / ref.abs().max()) < TOL
    dy = torch.randn(M, N, device="cuda") * 0.01
    torch.backends.cuda.matmul.allow_tf32 = False
    ref_dx, ref_dw = dy @ W, dy.t() @ x
    torch.backends.cuda.matmul.allow_tf32 = True
    assert float((ops.gemm(dy, W, b_mn=True) - ref_dx).abs().max() / ref_dx.abs().max()) < TOL
    assert float((ops.gemm(dy, x, a_mn=True, b_mn=True, splits=9) - ref_dw).abs().max() / ref_dw.abs().max()) < TOL


def test_rejects_bad_arguments():
    from pixelrec_b200 import ops
    from pixelrec_b200.lib import PixelRecB200Error
    x, W = torch.randn(64, 50, device="cuda"), torch.randn(32, 50, device="cuda")
    with pytest.raises(PixelRecB200Error):
        ops.gemm(x, W)                                    # K-major operands need K % 4 == 0 (16-byte row pitch for TMA)
    x, W = torch.randn(64, 48, device="cuda"), torch.randn(32, 48, device="cuda")     # a K tail (48 = 32 + 16) is fine
    ref = x.double() @ W.double().t()
    assert float((ops.gemm(x, W).double() - ref).abs().max() / ref.abs().max()) < TOL
    with pytest.raises(PixelRecB200Error):
        ops.gemm(torch.randn(64, 64), torch.randn(32, 64))   # CPU tensors: no fallback


def test_truncation_debias_removes_the_systematic_shrink():
    """The tensor core truncates fp32 operands to TF32: without compensation every output is ~0.07 % too small (a bias that
    compounds through a chain of GEMMs); PR_GEMM_DEBIAS makes the error zero-mean."""
    from pixelrec_b200 import ops
    g = np.random.default_rng(3)
    M, N, K = 512, 256, 1024
    x = np.abs(g.standard_normal((M, K))).astype(np.float32) + 0.5        # same-sign terms: the shrink shows directly
    W = np.abs(g.standard_normal((N, K))).astype(np.float32) + 0.5
    ref = x.astype(np.float64) @ W.astype(np.float64).T
    raw = ops.gemm(t(x), t(W)).double().cpu().numpy()
    fix = ops.gemm(t(x), t(W), debias=True).double().cpu().numpy()
    shrink = float((raw / ref - 1).mean())
    assert -0.9e-3 < shrink < -0.55e-3, shrink                       # ~2 * 0.72 * 2^-11 = -7.0e-4 (mantissa-averaged truncation)
    assert abs(float((fix / ref - 1).mean())) < 5e-5
    assert np.abs(fix / ref - 1).max() < 2e-4


@pytest.mark.parametrize("M,N,K", [(640, 512, 256), (300, 260, 96), (1000, 1536, 512), (81, 512, 1024), (4099, 1024, 64)])
def test_drop_add_epilogue_exact_and_mask_is_the_layernorm_mask(M, N, K):
    """pr_gemm_tf32_drop: out = dropout(x W^T + b) + res with the keep bits of pr_add_ln_fwd_f32 (oracle/philox_np.py
    rowwise_keep_scale).  Integer operands and p = 0.5 (scale 2) make every value exact: a wrong Philox counter, nibble or
    column mapping shows as a wrong integer."""
    from oracle import philox_np as PH
    from pixelrec_b200 import ops
    g = np.random.default_rng(M + N)
    x = g.integers(-3, 4, size=(M, K)).astype(np.float32)
    W = g.integers(-3, 4, size=(N, K)).astype(np.float32)
    b = g.integers(-5, 6, size=N).astype(np.float32)
    res = g.integers(-9, 10, size=(M, N)).astype(np.float32)
    keep = PH.rowwise_keep_scale(M, N, 0.5, 77, 5).astype(np.float64)          # 0 or 2
    ref = (x.astype(np.float64) @ W.astype(np.float64).T + b) * keep + res
    got = ops.gemm_drop_add(t(x), t(W), t(b), t(res), 0.5, 77, 5, debias=False).cpu().numpy().astype(np.float64)
    assert 0.4 < (keep == 0).mean() < 0.6
    assert np.array_equal(got, ref), f"{(got != ref).mean():.4f} of the outputs differ"
    # p = 0: plain bias + residual epilogue
    got0 = ops.gemm_drop_add(t(x), t(W), t(b), t(res), 0.0, 77, 5, debias=False).cpu().numpy().astype(np.float64)
    assert np.array_equal(got0, x.astype(np.float64) @ W.astype(np.float64).T + b + res)
    # another stream id draws another mask
    other = ops.gemm_drop_add(t(x), t(W), t(b), t(res), 0.5, 77, 6, debias=False).cpu().numpy()
    assert (other != got).mean() > 0.2
