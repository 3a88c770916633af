"""GPU parity: fused (dropout)+add+LayerNorm(+dropout), activations, and the pairwise loss vs the oracle.
fp32 tolerance 1e-5 relative-to-max (well inside north_star's 1e-3)."""
import numpy as np
import pytest
import torch

from oracle import philox_np as PH
from oracle import sasrec_np as O
from tests.gpu_util import dev, rel, t

pytestmark = pytest.mark.gpu
TOL = 2e-5


@pytest.fixture(params=[0, 1, 2, 5, 513], ids=["tune0", "tune_pipe", "tune_l2pf", "tune_pipe_fwd_bwd", "tune_fwd_rows2"])
def ln_tune(request):
    """runs a test under each LayerNorm kernel variant (pr_set_tuning); results must not depend on it"""
    from pixelrec_b200 import lib
    L_ = lib.load()
    before = L_.pr_set_tuning(-1)
    L_.pr_set_tuning(request.param)
    yield request.param
    L_.pr_set_tuning(before)


def _ln_ref(h, res, gamma, beta, eps, mpre, mpost):
    z = h * mpre + res
    y, cache = O.layernorm_fwd(z.astype(np.float64), gamma.astype(np.float64), beta.astype(np.float64), eps)
    return y * mpost, cache


@pytest.mark.parametrize("rows,D", [(1, 4), (37, 64), (100, 128), (333, 512), (17, 768), (65, 1024), (40, 2048), (9, 4096), (50, 36)])
@pytest.mark.parametrize("p_pre,p_post", [(0.0, 0.0), (0.1, 0.0), (0.0, 0.5)])
def test_add_ln_fwd_bwd(rows, D, p_pre, p_post, ln_tune):
    from pixelrec_b200 import ops
    g = np.random.default_rng(rows * D)
    h = g.standard_normal((rows, D)).astype(np.float32)
    res = g.standard_normal((rows, D)).astype(np.float32)
    gamma = (1 + 0.1 * g.standard_normal(D)).astype(np.float32)
    beta = (0.1 * g.standard_normal(D)).astype(np.float32)
    dy = g.standard_normal((rows, D)).astype(np.float32)
    seed, s_pre, s_post = 1234567, 3, 9
    mpre = PH.rowwise_keep_scale(rows, D, p_pre, seed, s_pre)
    mpost = PH.rowwise_keep_scale(rows, D, p_post, seed, s_post)
    y_ref, cache = _ln_ref(h, res, gamma, beta, 1e-12, mpre, mpost)
    dz, dg, db = O.layernorm_bwd((dy * mpost).astype(np.float64), gamma.astype(np.float64), cache)

    th, tr = t(h).requires_grad_(), t(res).requires_grad_()
    tg, tb = t(gamma).requires_grad_(), t(beta).requires_grad_()
    y = ops.add_ln(th, tr, tg, tb, 1e-12, p_pre=p_pre, p_post=p_post, seed=seed, stream_pre=s_pre, stream_post=s_post)
    y.backward(t(dy))
    torch.cuda.synchronize()
    assert rel(y.detach().cpu().numpy(), y_ref) < TOL
    if p_pre > 0 or p_post > 0:   # the mask really is the Philox one (exact zeros in the same places)
        zero_ref = (mpost == 0) if p_post > 0 else None
        if zero_ref is not None:
            assert np.array_equal(y.detach().cpu().numpy() == 0, zero_ref)
    assert rel(tr.grad.cpu().numpy(), dz) < TOL
    assert rel(th.grad.cpu().numpy(), dz * mpre) < TOL
    assert rel(tg.grad.cpu().numpy(), dg) < TOL and rel(tb.grad.cpu().numpy(), db) < TOL


@pytest.mark.parametrize("B,L,D", [(3, 10, 128), (5, 20, 512), (2, 7, 64)])
@pytest.mark.parametrize("p", [0.0, 0.25])
def test_embed_ln_strided_layout_and_posemb(B, L, D, p, ln_tune):
    """sasrec.py:72,77-83: LN(E[:,0,:-1] + P[0:L]) read in place from the [B,2,L+1,D] gather output;
    grad wrt E lands only on rows (b,0,t<L); grad wrt P is the sum over the batch."""
    from pixelrec_b200 import ops
    g = np.random.default_rng(B + L + D)
    E = g.standard_normal((B, 2, L + 1, D)).astype(np.float32)
    P = g.standard_normal((L + 3, D)).astype(np.float32)          # max_seq_length > L: extra rows get zero grad
    gamma = (1 + 0.1 * g.standard_normal(D)).astype(np.float32)
    beta = (0.1 * g.standard_normal(D)).astype(np.float32)
    dy = g.standard_normal((B, L, D)).astype(np.float32)
    mpost = PH.rowwise_keep_scale(B * L, D, p, 77, 0).reshape(B, L, D)
    z = E[:, 0, :-1] + P[:L][None]
    y_ref, cache = O.layernorm_fwd(z.astype(np.float64), gamma.astype(np.float64), beta.astype(np.float64), 1e-12)
    y_ref = y_ref * mpost
    dz, dg, db = O.layernorm_bwd((dy * mpost).astype(np.float64), gamma.astype(np.float64), cache)
    tE, tP = t(E).requires_grad_(), t(P).requires_grad_()
    tg, tb = t(gamma).requires_grad_(), t(beta).requires_grad_()
    y = ops.add_ln(tE, tP, tg, tb, 1e-12, p_post=p, seed=77, stream_post=0, layout=(L, 2 * (L + 1) * D, B), res_period=L)
    assert y.shape == (B, L, D)
    y.backward(t(dy))
    torch.cuda.synchronize()
    assert rel(y.detach().cpu().numpy(), y_ref) < TOL
    gE = tE.grad.cpu().numpy()
    assert rel(gE[:, 0, :-1], dz) < TOL
    assert (gE[:, 1] == 0).all() and (gE[:, 0, -1] == 0).all()
    gP = tP.grad.cpu().numpy()
    assert rel(gP[:L], dz.sum(0)) < TOL and (gP[L:] == 0).all()
    assert rel(tg.grad.cpu().numpy(), dg) < TOL and rel(tb.grad.cpu().numpy(), db) < TOL


@pytest.mark.parametrize("act", ["gelu", "relu", "swish", "tanh", "sigmoid"])
@pytest.mark.parametrize("n", [5, 4096, 100001])
def test_activation(act, n):
    from pixelrec_b200 import ops
    g = np.random.default_rng(n)
    x = (3 * g.standard_normal(n)).astype(np.float32)
    dy = g.standard_normal(n).astype(np.float32)
    x64 = x.astype(np.float64)
    sig = 1 / (1 + np.exp(-x64))
    ref = {"gelu": (O.gelu_fwd(x64), O.gelu_bwd(x64, dy.astype(np.float64))),
           "relu": (np.maximum(x64, 0), dy * (x64 > 0)),
           "swish": (x64 * sig, dy * (sig + x64 * sig * (1 - sig))),
           "tanh": (np.tanh(x64), dy * (1 - np.tanh(x64) ** 2)),
           "sigmoid": (sig, dy * sig * (1 - sig))}[act]
    tx = t(x).requires_grad_()
    y = ops.activation(tx, act)
    y.backward(t(dy))
    assert rel(y.detach().cpu().numpy(), ref[0]) < 1e-6 and rel(tx.grad.cpu().numpy(), ref[1]) < 1e-6
    with pytest.raises(KeyError):
        ops.activation(tx, "mish")


@pytest.mark.parametrize("B,L,D", [(4, 10, 128), (64, 20, 512), (3, 7, 64), (2, 20, 2048), (5, 3, 36), (3, 5, 256), (2, 4, 1024), (4, 6, 768)])
def test_bpr_loss_fwd_bwd(B, L, D):
    from pixelrec_b200 import ops
    g = np.random.default_rng(B * L + D)
    out = g.standard_normal((B, L, D)).astype(np.float32)
    E = (0.3 * g.standard_normal((B, 2, L + 1, D))).astype(np.float32)
    mask = (g.random((B, L)) < 0.7).astype(np.int64)
    mask[0] = 1
    mask[-1] = 0                                              # a fully padded sequence contributes exactly 0
    o64, e64 = out.astype(np.float64), E.astype(np.float64)
    loss_ref, cache = O.bpr_loss_fwd(o64, e64[:, 0, 1:], e64[:, 1, 1:], mask)
    d_out, d_tp, d_tn = O.bpr_loss_bwd(o64, e64[:, 0, 1:], e64[:, 1, 1:], cache, dloss=1.7)
    to, tE = t(out).requires_grad_(), t(E).requires_grad_()
    loss = ops.bpr_loss(to, tE, t(mask))
    (loss * 1.7).backward()
    torch.cuda.synchronize()
    assert abs(loss.item() - loss_ref) / abs(loss_ref) < TOL
    assert rel(to.grad.cpu().numpy(), d_out) < TOL
    gE = tE.grad.cpu().numpy()
    assert rel(gE[:, 0, 1:], d_tp) < TOL and rel(gE[:, 1, 1:], d_tn) < TOL
    assert (gE[:, :, 0] == 0).all()
    assert (gE[-1] == 0).all() and (to.grad[-1] == 0).all()


def test_bpr_loss_known_answer_at_zero_scores():
    """out == 0 -> sigmoid(0) = 0.5 -> loss = (#valid positions / B) * -log(0.5 + 1e-8)  (SURVEY 8c sanity)."""
    from pixelrec_b200 import ops
    B, L, D = 8, 10, 128
    out = torch.zeros(B, L, D, device=dev())
    E = torch.randn(B, 2, L + 1, D, device=dev())
    mask = torch.ones(B, L, dtype=torch.int64, device=dev())
    mask[:, :3] = 0
    loss = ops.bpr_loss(out, E, mask)
    assert abs(loss.item() - 7 * -np.log(0.5 + 1e-8)) < 1e-5


def test_bpr_loss_saturation_no_nan():
    from pixelrec_b200 import ops
    B, L, D = 2, 4, 64
    out = torch.full((B, L, D), 10.0, device=dev())
    E = torch.zeros(B, 2, L + 1, D, device=dev())
    E[:, 1] = 10.0                                            # pos - neg = -6400 -> sigmoid underflows to 0
    out.requires_grad_()
    loss = ops.bpr_loss(out, E, torch.ones(B, L, dtype=torch.int64, device=dev()))
    loss.backward()
    assert torch.isfinite(loss) and abs(loss.item() - L * -np.log(1e-8)) < 1e-3    # the reference's +1e-8 floor
    assert torch.isfinite(out.grad).all()


@pytest.mark.parametrize("rows,D,p", [(333, 512, 0.1), (100, 128, 0.0), (65, 1024, 0.1),
                                      # long enough that every warp of the pipelined variant refills its stages several times
                                      (12001, 512, 0.1), (40000, 128, 0.1), (30000, 256, 0.0), (9000, 1024, 0.1)])
def test_add_ln_bwd_with_bias_grad_partials(rows, D, p, ln_tune):
    """pr_add_ln_bwd_bias_f32: the extra partial matrix is the column sum of dh (bias grad of the producing Linear)."""
    from pixelrec_b200 import ops
    g = np.random.default_rng(rows + D)
    h = g.standard_normal((rows, D)).astype(np.float32)
    res = g.standard_normal((rows, D)).astype(np.float32)
    gamma = (1 + 0.1 * g.standard_normal(D)).astype(np.float32)
    beta = (0.1 * g.standard_normal(D)).astype(np.float32)
    dy = g.standard_normal((rows, D)).astype(np.float32)
    mpre = PH.rowwise_keep_scale(rows, D, p, 42, 7)
    y_ref, cache = _ln_ref(h, res, gamma, beta, 1e-12, mpre, np.ones_like(mpre))
    dz, dg, db = O.layernorm_bwd(dy.astype(np.float64), gamma.astype(np.float64), cache)
    y, mean, rstd = ops._raw_add_ln_fwd(t(h), t(res), t(gamma), t(beta), 1e-12, p, 42, 7)
    assert rel(y.cpu().numpy(), y_ref) < TOL
    dh, dres, dgam, dbet, dbias = ops._raw_add_ln_bwd_bias(t(dy), t(h), t(res), t(gamma), mean, rstd, p, 42, 7)
    assert rel(dh.cpu().numpy(), dz * mpre) < TOL and rel(dres.cpu().numpy(), dz) < TOL
    assert rel(dgam.cpu().numpy(), dg) < TOL and rel(dbet.cpu().numpy(), db) < TOL
    assert rel(dbias.cpu().numpy(), (dz * mpre).sum(0)) < TOL
    if ln_tune:   # same arithmetic per row under every variant: dh / dres are bit-identical to the register kernel
        from pixelrec_b200 import lib
        lib.load().pr_set_tuning(0)
        dh0, dres0, *_ = ops._raw_add_ln_bwd_bias(t(dy), t(h), t(res), t(gamma), mean, rstd, p, 42, 7)
        assert torch.equal(dh0, dh) and torch.equal(dres0, dres)


@pytest.mark.parametrize("rows,D,p", [(333, 512, 0.1), (100, 128, 0.0), (65, 1024, 0.1), (12001, 512, 0.1), (40000, 128, 0.1),
                                      (30000, 256, 0.0), (9000, 1024, 0.1), (500, 2048, 0.1), (77, 36, 0.1)])
def test_ln_backward_from_stored_z(rows, D, p, ln_tune):
    """pr_add_ln_bwd_bias_z_f32: the producing GEMM stored z = drop(h) + res, so LayerNorm runs on z alone (forward: res = NULL,
    p_pre = 0) and the backward reads dy and z; dh still carries the dropout mask of (seed, stream), dz does not."""
    from pixelrec_b200 import ops
    g = np.random.default_rng(rows + D)
    h = g.standard_normal((rows, D)).astype(np.float32)
    res = g.standard_normal((rows, D)).astype(np.float32)
    gamma = (1 + 0.1 * g.standard_normal(D)).astype(np.float32)
    beta = (0.1 * g.standard_normal(D)).astype(np.float32)
    dy = g.standard_normal((rows, D)).astype(np.float32)
    mpre = PH.rowwise_keep_scale(rows, D, p, 42, 7)
    z = (h * mpre + res).astype(np.float32)
    y_ref, cache = _ln_ref(z, np.zeros_like(z), gamma, beta, 1e-12, np.ones_like(mpre), np.ones_like(mpre))
    dz_ref, dg, db = O.layernorm_bwd(dy.astype(np.float64), gamma.astype(np.float64), cache)
    y, mean, rstd = ops._raw_add_ln_fwd(t(z), None, t(gamma), t(beta), 1e-12, 0.0, 42, 7)
    assert rel(y.cpu().numpy(), y_ref) < TOL
    dh, dz, dgam, dbet, dbias = ops._raw_ln_z_bwd_bias(t(dy), t(z), t(gamma), mean, rstd, p, 42, 7)
    assert rel(dh.cpu().numpy(), dz_ref * mpre) < TOL and rel(dz.cpu().numpy(), dz_ref) < TOL
    assert rel(dgam.cpu().numpy(), dg) < TOL and rel(dbet.cpu().numpy(), db) < TOL
    assert rel(dbias.cpu().numpy(), (dz_ref * mpre).sum(0)) < TOL
    if p == 0.0:
        assert dz.data_ptr() == dh.data_ptr()                  # one tensor when there is no mask


def test_fused_layer_z_mode_equals_separate_kernels():
    """TransformerLayerFn with dense / dense_2 writing z = dropout(linear) + residual from the GEMM epilogue (PR_FUSE_LN_Z, default)
    == the same layer with the separate add+LayerNorm kernels: same Philox draws, same arithmetic up to fused multiply-adds."""
    import pixelrec_b200.model.layers as Lm
    from pixelrec_b200 import ops
    tf32_before = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True            # the epilogue fusion belongs to the tensor-core GEMM path
    torch.manual_seed(1)
    layer = Lm.TransformerLayer(4, 512, 1024, 0.1, 0.1, "gelu", 1e-12).to(dev()).train()
    for p_ in layer.parameters():
        if p_.ndim == 1:
            p_.data.add_(0.05 * torch.randn_like(p_))
    x = torch.randn(40, 20, 512, device=dev())
    ids = torch.ones(40, 20, dtype=torch.int64, device=dev())
    ids[0, :5] = 0
    dyv = torch.randn(40, 20, 512, device=dev())
    outs = {}
    old = ops.FUSE_LN_Z
    try:
        for zmode in (True, False):
            ops.FUSE_LN_Z = zmode
            layer.zero_grad()
            xi = x.clone().requires_grad_()
            y = layer(xi, ids, True, 123, 1)
            y.backward(dyv)
            outs[zmode] = (y.detach(), xi.grad.clone(), {k: v.grad.clone() for k, v in layer.named_parameters()})
    finally:
        ops.FUSE_LN_Z = old
        torch.backends.cuda.matmul.allow_tf32 = tf32_before
    assert torch.allclose(outs[True][0], outs[False][0], rtol=1e-5, atol=2e-6)
    assert (outs[True][1] - outs[False][1]).abs().max().item() <= 1e-4 * outs[False][1].abs().max().item()
    for k in outs[True][2]:
        a, b = outs[True][2][k], outs[False][2][k]
        assert (a - b).abs().max().item() <= 1e-4 * max(b.abs().max().item(), 1e-3) + 1e-6, k


@pytest.mark.parametrize("rows,cols,act", [(500, 1024, "gelu"), (33, 256, "relu"), (7, 4096, "gelu"), (129, 36, "swish")])
def test_act_bwd_with_bias_grad(rows, cols, act):
    from pixelrec_b200 import ops
    g = np.random.default_rng(rows)
    x = (2 * g.standard_normal((rows, cols))).astype(np.float32)
    dy = g.standard_normal((rows, cols)).astype(np.float32)
    tx = t(x).requires_grad_()
    ops.activation(tx, act).backward(t(dy))
    dx, db = ops._raw_act_bwd_bias(t(x), t(dy), ops.ACT_IDS[act])
    assert rel(dx.cpu().numpy(), tx.grad.cpu().numpy()) < 1e-6
    assert rel(db.cpu().numpy(), tx.grad.cpu().numpy().astype(np.float64).sum(0)) < 2e-5


def test_fused_layer_equals_op_by_op_layer():
    """ops.TransformerLayerFn (hand-written backward) == the op-by-op composition of the same kernels."""
    import pixelrec_b200.model.layers as Lm
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    layer = Lm.TransformerLayer(4, 128, 256, 0.1, 0.1, "gelu", 1e-12).to(dev()).train()
    for p_ in layer.parameters():
        if p_.ndim == 1:
            p_.data.add_(0.05 * torch.randn_like(p_))
    x = torch.randn(6, 20, 128, device=dev())
    ids = torch.ones(6, 20, dtype=torch.int64, device=dev())
    ids[0, :5] = 0
    dyv = torch.randn(6, 20, 128, device=dev())
    outs = {}
    for fused in (True, False):
        Lm.FUSED_LAYER = fused
        layer.zero_grad()
        xi = x.clone().requires_grad_()
        y = layer(xi, ids, True, 123, 1)
        y.backward(dyv)
        outs[fused] = (y.detach(), xi.grad.clone(), {k: v.grad.clone() for k, v in layer.named_parameters()})
    Lm.FUSED_LAYER = True
    assert torch.allclose(outs[True][0], outs[False][0], rtol=1e-5, atol=1e-6)
    assert torch.allclose(outs[True][1], outs[False][1], rtol=1e-4, atol=1e-6)
    for k in outs[True][2]:
        a, b = outs[True][2][k], outs[False][2][k]
        # + 1e-6: gradients that are zero in exact arithmetic (the key bias: softmax is shift-invariant) are pure rounding noise
        assert (a - b).abs().max().item() <= 1e-4 * max(b.abs().max().item(), 1e-3) + 1e-6, k
    torch.backends.cuda.matmul.allow_tf32 = True


def test_weight_grad_split_k_matches_single_gemm():
    """ops._wgrad: the explicit split-K form (used when M = B*L is large) equals dy^T @ x."""
    from pixelrec_b200 import ops
    g = np.random.default_rng(5)
    M, o, i = 8 * 2048, 96, 64
    dy = g.standard_normal((M, o)).astype(np.float32)
    x = g.standard_normal((M, i)).astype(np.float32)
    ref = dy.astype(np.float64).T @ x.astype(np.float64)
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        assert ops.WGRAD_SPLIT > 1 and M % ops.WGRAD_SPLIT == 0 and M // ops.WGRAD_SPLIT >= 2048
        got = ops._wgrad(t(dy), t(x))
        small = ops._wgrad(t(dy[:1000]), t(x[:1000]))          # below the threshold: plain GEMM
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    assert got.shape == (o, i) and rel(got.cpu().numpy(), ref) < TOL
    assert rel(small.cpu().numpy(), dy[:1000].astype(np.float64).T @ x[:1000].astype(np.float64)) < TOL


@pytest.mark.parametrize("M,C", [(1, 4), (70, 16), (1000, 132), (81920, 1536), (4097, 512)])
def test_colsum_rows(M, C):
    """bias gradient of the fused q|k|v projection: column sums of its output gradient (layers.py:586-588)"""
    from pixelrec_b200 import ops
    g = np.random.default_rng(M + C)
    x = g.standard_normal((M, C)).astype(np.float32)
    got = ops.colsum_rows(t(x))
    ref = x.astype(np.float64).sum(0)
    assert got.shape == (C,)
    assert np.abs(got.cpu().numpy() - ref).max() <= 1e-5 * np.abs(x).sum(0).max()
    assert torch.equal(got, ops.colsum_rows(t(x)))          # deterministic
