"""Tensor-level wrappers and autograd Functions over the C ABI (pixelrec_b200/lib.py).

PyTorch is plumbing here: device memory, the current stream, autograd bookkeeping and the cuBLAS
linear layers.  Every function below launches hand-written sm_100a kernels through
libpixelrec_b200.so and raises if given anything but CUDA tensors -- there is no CPU / eager fallback.
"""
from __future__ import annotations

import math
import os

import torch

from . import lib as _lib

ACT_IDS = {"gelu": 0, "relu": 1, "swish": 2, "tanh": 3, "sigmoid": 4, "quick_gelu": 5}
LAUNCHES = {"count": 0}   # kernels of OURS launched (bench.py reports it as gpu_launches)

_cur_dev = [None]


def _L():
    return _lib.load()


def _p(t):
    return None if t is None else t.data_ptr()


def _stream(t):
    dev = t.device.index
    if _cur_dev[0] != dev:
        _lib.check(_L().pr_set_device(dev), "pr_set_device")
        _cur_dev[0] = dev
    return torch.cuda.current_stream(t.device).cuda_stream


def _req(t, dtype, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.PixelRecB200Error(f"{name}: expected a CUDA tensor (pixelrec_b200 has no CPU fallback), got "
                                     f"{type(t).__name__} on {getattr(t, 'device', None)}")
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name}: tensor must be contiguous")
    return t


def _count(n=1):
    LAUNCHES["count"] += n


# Optional live kernel timing (bench.py roofline): PROFILE["names"] = set of kernel names (or None = all);
# every wrapped launch then records a CUDA event pair on the launching stream into PROFILE["events"][name].
PROFILE = {"on": False, "names": None, "events": {}}


class _prof:
    __slots__ = ("name", "dev", "s")

    def __init__(self, name, t):
        self.name = name
        self.dev = t.device
        self.s = None

    def __enter__(self):
        if PROFILE["on"] and (PROFILE["names"] is None or self.name in PROFILE["names"]):
            self.s = torch.cuda.Event(enable_timing=True)
            self.s.record(torch.cuda.current_stream(self.dev))
        return self

    def __exit__(self, *exc):
        if self.s is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record(torch.cuda.current_stream(self.dev))
            PROFILE["events"].setdefault(self.name, []).append((self.s, e))
        return False


def profile_summary():
    """name -> (n_launches, mean_ms); call after torch.cuda.synchronize()."""
    out = {}
    for name, evs in PROFILE["events"].items():
        ms = [s.elapsed_time(e) for s, e in evs]
        out[name] = (len(ms), sum(ms) / max(len(ms), 1))
    return out


# ------------------------------------------------------------------------------------------- K1 gather
def gather_rows(W, idx, impl=0, status=None):
    """out[..., :] = W[idx[...], :]  (REC/model/IDNet/sasrec.py:68).  impl: 0 auto, 1 LDG, 2 TMA bulk."""
    _req(W, torch.float32, "W")
    _req(idx, torch.int64, "idx")
    N, D = W.shape
    R = idx.numel()
    out = torch.empty(*idx.shape, D, device=W.device, dtype=torch.float32)
    with _prof("gather_rows", W):
        _lib.check(_L().pr_gather_rows_f32(_p(W), N, D, _p(idx), R, _p(out), _p(status), impl, _stream(W)), "pr_gather_rows_f32")
    _count()
    return out


# ------------------------------------------------------------------------------------------- K2 scatter
class ScatterPlan:
    """Device-side sort/segment plan of one index tensor (pr_scatter_plan)."""

    def __init__(self, idx, N, padding_idx=0, row2slot=None, status=None):
        _req(idx, torch.int64, "idx")
        dev = idx.device
        self.R = idx.numel()
        self.N = int(N)
        self.max_uniq = max(1, min(self.R, self.N))
        i32 = dict(device=dev, dtype=torch.int32)
        self.perm = torch.empty(max(self.R, 1), **i32)
        self.uniq_ids = torch.empty(self.max_uniq, **i32)
        self.seg_start = torch.empty(self.max_uniq + 1, **i32)
        self.n_uniq = torch.empty(1, **i32)
        self.row2slot = row2slot
        ws_bytes = _L().pr_scatter_plan_workspace_bytes(self.R, self.N)
        self._ws = torch.empty(max(ws_bytes, 16), device=dev, dtype=torch.uint8)
        pad = -1 if padding_idx is None else int(padding_idx)
        with _prof("scatter_plan", idx):
            _lib.check(_L().pr_scatter_plan(_p(idx), self.R, self.N, pad, _p(self.perm), _p(self.uniq_ids),
                                            _p(self.seg_start), _p(self.n_uniq), _p(row2slot), _p(self._ws), ws_bytes,
                                            _p(status), _stream(idx)), "pr_scatter_plan")
        bits = max(1, int(self.N).bit_length())
        _count(3 * ((bits + 9) // 10) + 3 if self.R else 1)      # digits of up to 10 bits; convert / flag / emit ride in other passes


def scatter_add_rows(dOut, plan: ScatterPlan, scale=1.0, dense_G=None, out_rows=True):
    """Sparse gradient rows [max_uniq, D] (rows >= n_uniq are undefined) and/or dense_G[uniq] = rows."""
    _req(dOut, torch.float32, "dOut")
    D = dOut.shape[-1]
    if dOut.numel() != plan.R * D:
        raise ValueError("scatter_add_rows: dOut does not match the plan")
    rows = torch.empty(plan.max_uniq, D, device=dOut.device, dtype=torch.float32) if out_rows else None
    if dense_G is not None:
        _req(dense_G, torch.float32, "dense_G")
    with _prof("scatter_add_rows", dOut):
        _lib.check(_L().pr_scatter_add_rows_f32(_p(dOut), plan.R, D, _p(plan.perm), _p(plan.uniq_ids), _p(plan.seg_start),
                                                _p(plan.n_uniq), plan.max_uniq, float(scale), _p(rows), _p(dense_G),
                                                _stream(dOut)), "pr_scatter_add_rows_f32")
    _count()
    return rows


# ------------------------------------------------------------------------------------------- K10 AdamW
def adamw_rows(W, M, V, grad_rows, row2slot, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0, step_dev=None):
    _req(W, torch.float32, "W"); _req(M, torch.float32, "M"); _req(V, torch.float32, "V")
    N, D = W.shape
    with _prof("adamw_rows", W):
        _lib.check(_L().pr_adamw_rows_f32(_p(W), _p(M), _p(V), N, D, _p(grad_rows), _p(row2slot), lr, beta1, beta2, eps,
                                          weight_decay, grad_scale, int(step), _p(step_dev), _stream(W)), "pr_adamw_rows_f32")
    _count()


def adamw_dense(w, g, m, v, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0, step_dev=None):
    for t, n in ((w, "w"), (g, "g"), (m, "m"), (v, "v")):
        _req(t, torch.float32, n)
    with _prof("adamw_dense", w):
        _lib.check(_L().pr_adamw_dense_f32(_p(w), _p(g), _p(m), _p(v), w.numel(), lr, beta1, beta2, eps, weight_decay,
                                           grad_scale, int(step), _p(step_dev), _stream(w)), "pr_adamw_dense_f32")
    _count()


# ------------------------------------------------------------------------------------------- dropout rng
class DropoutRng:
    """Philox seed/stream bookkeeping.  One seed per forward call, one stream id per dropout site."""

    def __init__(self, seed=0):
        self.base = int(seed) & 0xFFFFFFFFFFFF
        self.calls = 0

    def next_seed(self):
        self.calls += 1
        return (self.base * 1000003 + self.calls) & 0xFFFFFFFFFFFFFFFF


def set_seed_device(offset):
    """offset: uint64-sized CUDA tensor (int64 [1]) added to every dropout seed by the kernels, or None for host seeds only.
    Raises PixelRecB200Error in builds without -DPR_SEED_DEV (include/pixelrec_b200.h)."""
    if offset is not None:
        _req(offset, torch.int64, "offset")
    _lib.check(_L().pr_set_seed_device(_p(offset)), "pr_set_seed_device")


# ------------------------------------------------------------------------------------------- K3/K7 add+LN
class GradSlab:
    """Shared [B,2,L+1,D] table-gradient buffer: the loss backward creates it, the embedding LayerNorm
    backward accumulates its rows in place, the gather backward consumes it (see SASRec.forward)."""

    def __init__(self):
        self.dE = None


class AddLnFn(torch.autograd.Function):
    """y = drop_post(LN(drop_pre(h) + res)).  `layout` = None for contiguous h [rows, D], or
    (rows_per_seq, seq_stride_floats, n_seq) to read rows (s,t) out of a larger tensor in place."""

    @staticmethod
    def forward(ctx, h, res, gamma, beta, eps, p_pre, p_post, seed, stream_pre, stream_post, layout, res_period, slab):
        _req(h, torch.float32, "h"); _req(gamma, torch.float32, "gamma"); _req(beta, torch.float32, "beta")
        if res is not None:
            _req(res, torch.float32, "res")
        D = gamma.numel()
        if layout is None:
            rows = h.numel() // D
            rps, sstride = rows if rows > 0 else 1, 0
            out_shape = h.shape
        else:
            rps, sstride, nseq = layout
            rows = rps * nseq
            out_shape = (nseq, rps, D)
        y = torch.empty(out_shape, device=h.device, dtype=torch.float32)
        mean = torch.empty(max(rows, 1), device=h.device, dtype=torch.float32)
        rstd = torch.empty(max(rows, 1), device=h.device, dtype=torch.float32)
        with _prof("add_ln_fwd", h):
            _lib.check(_L().pr_add_ln_fwd_f32(_p(h), sstride, rps, _p(res), res_period, _p(gamma), _p(beta), eps, rows, D,
                                              p_pre, p_post, seed, stream_pre, stream_post, _p(y), _p(mean), _p(rstd),
                                              _stream(h)), "pr_add_ln_fwd_f32")
        _count()
        ctx.save_for_backward(h, res, gamma, mean, rstd)
        ctx.cfg = (eps, p_pre, p_post, seed, stream_pre, stream_post, layout, res_period, rows, D, rps, sstride)
        ctx.slab = slab
        return y

    @staticmethod
    def backward(ctx, dy):
        h, res, gamma, mean, rstd = ctx.saved_tensors
        eps, p_pre, p_post, seed, s_pre, s_post, layout, res_period, rows, D, rps, sstride = ctx.cfg
        dy = dy.contiguous()
        n_part = _L().pr_add_ln_bwd_partials(rows, D)
        partials = torch.empty(2, n_part, D, device=dy.device, dtype=torch.float32)
        need_res = res is not None and ctx.needs_input_grad[1]
        # without pre-add dropout dh == dz == dres: one buffer serves both grads
        share = need_res and res_period <= 0 and p_pre == 0.0 and layout is None
        dres_rows = torch.empty(rows, D, device=dy.device, dtype=torch.float32) if (need_res and res_period <= 0 and not share) else None
        slab = ctx.slab
        ret_dh = None
        if layout is None:
            dh = torch.empty_like(h)
            dh_stride, acc = 0, 0
            ret_dh = dh
        elif slab is not None and slab.dE is not None:
            dh, dh_stride, acc = slab.dE, sstride, 1           # accumulate into the shared table-grad slab
        else:
            dh = torch.zeros_like(h)
            dh_stride, acc = sstride, 0
            ret_dh = dh
        # position-embedding style residual (broadcast over sequences): its grad is the per-t sum of dz
        dz_tmp = None
        if need_res and res_period > 0:
            dz_tmp = torch.empty(rows, D, device=dy.device, dtype=torch.float32)
        with _prof("add_ln_bwd", dy):
            _lib.check(_L().pr_add_ln_bwd_f32(_p(dy), _p(h), sstride, rps, _p(res), res_period, _p(gamma), _p(mean), _p(rstd),
                                              rows, D, p_pre, p_post, seed, s_pre, s_post, _p(dh), dh_stride, acc,
                                              _p(dres_rows if dres_rows is not None else dz_tmp), _p(partials), n_part,
                                              _stream(dy)), "pr_add_ln_bwd_f32")
        dgb = torch.empty(2, D, device=dy.device, dtype=torch.float32)
        with _prof("colsum", dy):
            _lib.check(_L().pr_colsum_f32(_p(partials), 2, n_part, D, _p(dgb), _stream(dy)), "pr_colsum_f32")
        _count(2)
        dres = None
        if need_res:
            if res_period > 0:
                dres = torch.zeros_like(res)
                dres[:res_period] = dz_tmp.view(-1, res_period, D).sum(0)
            elif share:
                dres = ret_dh.view(res.shape)
            else:
                dres = dres_rows.view(res.shape)
        return ret_dh, dres, dgb[0], dgb[1], None, None, None, None, None, None, None, None, None


def add_ln(h, res, gamma, beta, eps, p_pre=0.0, p_post=0.0, seed=0, stream_pre=0, stream_post=0, layout=None,
           res_period=0, slab=None):
    return AddLnFn.apply(h, res, gamma, beta, float(eps), float(p_pre), float(p_post), int(seed), int(stream_pre),
                         int(stream_post), layout, int(res_period), slab)


# ------------------------------------------------------------------------------------------- activation
class ActFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, act):
        _req(x, torch.float32, "x")
        y = torch.empty_like(x)
        with _prof("act_fwd", x):
            _lib.check(_L().pr_act_fwd_f32(_p(x), x.numel(), act, _p(y), _stream(x)), "pr_act_fwd_f32")
        _count()
        ctx.save_for_backward(x)
        ctx.act = act
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        dy = dy.contiguous()
        dx = torch.empty_like(x)
        with _prof("act_bwd", x):
            _lib.check(_L().pr_act_bwd_f32(_p(x), _p(dy), x.numel(), ctx.act, _p(dx), _stream(x)), "pr_act_bwd_f32")
        _count()
        return dx, None


def activation(x, name):
    if name not in ACT_IDS:
        raise KeyError(f"hidden_act {name!r} not in {sorted(ACT_IDS)} (REC/model/layers.py:640-648)")
    return ActFn.apply(x, ACT_IDS[name])


# ------------------------------------------------------------------------------------------- K4+K6 attention
class AttnFn(torch.autograd.Function):
    """ctx = softmax(q k^T / sqrt(dh) + mask(key_ids, causal)) v on a fused [B, L, 3*D] q|k|v tensor."""

    @staticmethod
    def forward(ctx, qkv, key_ids, n_heads, causal, p_drop, seed, rng_stream, tf32=None):
        _req(qkv, torch.float32, "qkv")
        B, Lq, D3 = qkv.shape
        D = D3 // 3
        dh = D // n_heads
        if tf32 is None:   # follow the policy of the linear layers (model.layers / matmul_precision)
            tf32 = torch.backends.cuda.matmul.allow_tf32
        tf32 = bool(tf32) and Lq <= 32 and dh % 32 == 0 and (dh <= 128 or dh % 128 == 0)
        fwd = _L().pr_sasrec_attn_fwd_tf32 if tf32 else _L().pr_sasrec_attn_fwd_f32
        if key_ids is not None:
            _req(key_ids, torch.int64, "key_ids")
        out = torch.empty(B, Lq, D, device=qkv.device, dtype=torch.float32)
        probs = torch.empty(B, n_heads, Lq, Lq, device=qkv.device, dtype=torch.float32)
        base = qkv.data_ptr()
        with _prof("attn_fwd", qkv):
            _lib.check(fwd(base, base + 4 * D, base + 8 * D, D3, _p(key_ids), B, Lq, n_heads, dh,
                           int(causal), p_drop, seed, rng_stream, _p(out), _p(probs), _stream(qkv)), "pr_sasrec_attn_fwd")
        _count()
        ctx.save_for_backward(qkv, probs)
        ctx.cfg = (B, Lq, n_heads, dh, int(causal), p_drop, seed, rng_stream, tf32)
        return out

    @staticmethod
    def backward(ctx, dctx):
        qkv, probs = ctx.saved_tensors
        B, Lq, h, dh, causal, p_drop, seed, rng_stream, tf32 = ctx.cfg
        D = h * dh
        dctx = dctx.contiguous()
        dqkv = torch.empty_like(qkv)
        base, gbase = qkv.data_ptr(), dqkv.data_ptr()
        bwd = _L().pr_sasrec_attn_bwd_tf32 if tf32 else _L().pr_sasrec_attn_bwd_f32
        with _prof("attn_bwd", qkv):
            _lib.check(bwd(base, base + 4 * D, base + 8 * D, 3 * D, _p(probs), _p(dctx), B, Lq, h, dh, causal, p_drop, seed,
                           rng_stream, gbase, gbase + 4 * D, gbase + 8 * D, 3 * D, _stream(qkv)), "pr_sasrec_attn_bwd")
        _count()
        return dqkv, None, None, None, None, None, None, None


class LongAttnFn(torch.autograd.Function):
    """Same contract as AttnFn for 64 < L <= 256 (ViT-B/16's 197 tokens): strict fp32, no dropout, probabilities are
    recomputed in the backward from the saved log-sum-exp (pr_attn_long_*_f32, csrc/attn_long.cuh)."""

    @staticmethod
    def forward(ctx, qkv, key_ids, n_heads, causal):
        _req(qkv, torch.float32, "qkv")
        B, Lq, D3 = qkv.shape
        D = D3 // 3
        dh = D // n_heads
        if key_ids is not None:
            _req(key_ids, torch.int64, "key_ids")
        out = torch.empty(B, Lq, D, device=qkv.device, dtype=torch.float32)
        lse = torch.empty(B * n_heads, Lq, device=qkv.device, dtype=torch.float32)
        base = qkv.data_ptr()
        with _prof("attn_long_fwd", qkv):
            _lib.check(_L().pr_attn_long_fwd_f32(base, base + 4 * D, base + 8 * D, D3, _p(key_ids), B, Lq, n_heads, dh,
                                                 int(causal), _p(out), _p(lse), _stream(qkv)), "pr_attn_long_fwd_f32")
        _count()
        ctx.save_for_backward(qkv, out, lse)
        ctx.key_ids = key_ids
        ctx.cfg = (B, Lq, n_heads, dh, int(causal))
        return out

    @staticmethod
    def backward(ctx, dctx):
        qkv, out, lse = ctx.saved_tensors
        B, Lq, h, dh, causal = ctx.cfg
        D = h * dh
        dctx = dctx.contiguous()
        dqkv = torch.empty_like(qkv)
        delta = torch.empty_like(lse)
        base, gbase = qkv.data_ptr(), dqkv.data_ptr()
        with _prof("attn_long_bwd", qkv):
            _lib.check(_L().pr_attn_long_bwd_f32(base, base + 4 * D, base + 8 * D, 3 * D, _p(ctx.key_ids), _p(out), _p(lse),
                                                 _p(dctx), B, Lq, h, dh, causal, gbase, gbase + 4 * D, gbase + 8 * D, 3 * D,
                                                 _p(delta), _stream(qkv)), "pr_attn_long_bwd_f32")
        _count(2)
        return dqkv, None, None, None


def attention(qkv, key_ids, n_heads, causal=True, p_drop=0.0, seed=0, rng_stream=0, tf32=None):
    """tf32: None = follow torch.backends.cuda.matmul.allow_tf32 (the linear layers' policy), True/False to force.
    Sequences longer than 64 (ViT-B/16 item encoder) take the long-sequence kernels, which have no dropout."""
    if qkv.dim() == 3 and qkv.shape[1] > 64:
        if p_drop:
            raise _lib.PixelRecB200Error(f"attention: dropout is not supported for L={qkv.shape[1]} > 64")
        return LongAttnFn.apply(qkv, key_ids, int(n_heads), bool(causal))
    want_tf32 = bool(torch.backends.cuda.matmul.allow_tf32) if tf32 is None else bool(tf32)
    if qkv.dim() == 3 and 32 < qkv.shape[1] <= 64 and not p_drop and want_tf32:
        # CLIP ViT-B/32's 50 tokens: beyond the tensor-core kernel for short sequences (L <= 32), which left them to the FFMA
        # kernel (0.36 ms per layer at 352 images); the long-sequence kernels have an mma.sync forward and no lower bound on L
        dh = qkv.shape[2] // 3 // int(n_heads)
        if dh in (4, 8, 16, 32, 64, 128):
            return LongAttnFn.apply(qkv, key_ids, int(n_heads), bool(causal))
    return AttnFn.apply(qkv, key_ids, int(n_heads), bool(causal), float(p_drop), int(seed), int(rng_stream), tf32)


# ------------------------------------------------------------------------------------------- K8 loss
class BprLossFn(torch.autograd.Function):
    """loss(out [B,L,D], E [B,2,L+1,D], mask [B,L]) of sasrec.py:88-92; targets are read in place from E."""

    @staticmethod
    def forward(ctx, out, E, mask, slab):
        _req(out, torch.float32, "out"); _req(E, torch.float32, "E"); _req(mask, torch.int64, "masked_index")
        B, L, D = out.shape
        if tuple(E.shape) != (B, 2, L + 1, D):
            raise ValueError(f"E must be [B,2,L+1,D]={B, 2, L + 1, D}, got {tuple(E.shape)}")
        coef = torch.empty(B, L, device=out.device, dtype=torch.float32)
        terms = torch.empty(B, L, device=out.device, dtype=torch.float32)
        scores = torch.empty(2, B, L, device=out.device, dtype=torch.float32)
        loss = torch.empty((), device=out.device, dtype=torch.float32)
        eb = E.data_ptr()
        plane = (L + 1) * D
        with _prof("bpr_fwd", out):
            _lib.check(_L().pr_bpr_loss_fwd_f32(_p(out), eb + 4 * D, eb + 4 * (plane + D), 2 * plane, _p(mask), B, L, D,
                                                _p(scores[0]), _p(scores[1]), _p(coef), _p(terms), _p(loss), _stream(out)),
                       "pr_bpr_loss_fwd_f32")
        _count(2)
        ctx.save_for_backward(out, E, coef)
        ctx.slab = slab
        ctx.scores = scores
        return loss

    @staticmethod
    def backward(ctx, dloss):
        out, E, coef = ctx.saved_tensors
        B, L, D = out.shape
        dloss = dloss.contiguous().to(torch.float32)
        d_out = torch.empty_like(out)
        dE = torch.empty_like(E)
        dE[:, :, 0].zero_()   # rows (b,0,0) receive only the embedding-LN grad; (b,1,0) is dead (sasrec.py:74)
        gb = dE.data_ptr()
        eb = E.data_ptr()
        plane = (L + 1) * D
        with _prof("bpr_bwd", out):
            _lib.check(_L().pr_bpr_loss_bwd_f32(_p(out), eb + 4 * D, eb + 4 * (plane + D), 2 * plane, _p(coef), _p(dloss),
                                                B, L, D, _p(d_out), gb + 4 * D, gb + 4 * (plane + D), 2 * plane,
                                                _stream(out)), "pr_bpr_loss_bwd_f32")
        _count()
        if ctx.slab is not None:
            ctx.slab.dE = dE
        return d_out, dE, None, None


def bpr_loss(out, E, masked_index, slab=None):
    return BprLossFn.apply(out, E, masked_index, slab)


# ------------------------------------------------------------------------------------------- table gather w/ sparse grad
PLAN_AHEAD = os.environ.get("PR_PLAN_AHEAD", "0") == "1"     # build the scatter plan during the forward, on a side stream (r02a: no gain)
_plan_streams = {}


def _plan_stream(device):
    s = _plan_streams.get(device)
    if s is None:
        s = _plan_streams[device] = torch.cuda.Stream(device=device)
    return s


class GatherFn(torch.autograd.Function):
    """E = W[idx].  backward: dense mode returns a zero-filled [N,D] grad (reference semantics);
    sparse mode deposits (plan, reduced rows) on `sink` and leaves W.grad untouched.
    PR_PLAN_AHEAD=1: the index plan of the backward (9 small sort / scan launches that depend only on idx) is enqueued on a
    side stream at forward time, ordered after everything already on the current stream (the previous step's AdamW hands the
    row2slot entries back), and overlaps the encoder; the backward only waits for its event."""

    @staticmethod
    def forward(ctx, W, idx, padding_idx, sink, impl):
        out = gather_rows(W, idx, impl)
        ctx.save_for_backward(idx)
        ctx.N = W.shape[0]
        ctx.padding_idx = padding_idx
        ctx.sink = sink
        ctx.plan = None
        if PLAN_AHEAD and sink is not None and sink.sparse and ctx.needs_input_grad[0]:
            cur = torch.cuda.current_stream(idx.device)
            side = _plan_stream(idx.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                ctx.plan = ScatterPlan(idx, ctx.N, padding_idx, row2slot=sink.row2slot)
                ctx.plan_ready = torch.cuda.Event()
                ctx.plan_ready.record(side)
        return out

    @staticmethod
    def backward(ctx, dE):
        (idx,) = ctx.saved_tensors
        dE = dE.contiguous()
        sink = ctx.sink
        if sink is not None and sink.sparse:
            plan = ctx.plan
            if plan is None:
                plan = ScatterPlan(idx, ctx.N, ctx.padding_idx, row2slot=sink.row2slot)
            else:
                cur = torch.cuda.current_stream(idx.device)
                cur.wait_event(ctx.plan_ready)
                for tns in (plan.perm, plan.uniq_ids, plan.seg_start, plan.n_uniq, plan._ws):   # allocated on the side stream
                    tns.record_stream(cur)
            rows = scatter_add_rows(dE, plan)
            sink.deposit(plan, rows)
            return None, None, None, None, None
        plan = ScatterPlan(idx, ctx.N, ctx.padding_idx)
        G = torch.zeros(ctx.N, dE.shape[-1], device=dE.device, dtype=torch.float32)
        scatter_add_rows(dE, plan, dense_G=G, out_rows=False)
        return G, None, None, None, None


# ------------------------------------------------------------------------------------------- fused transformer layer
# One autograd node per encoder layer (REC/model/layers.py:676-703).  Same kernels and the same cuBLAS GEMMs as the
# op-by-op path above, but the backward is written out by hand, which removes what autograd adds around them:
#   * the two residual-gradient adds per layer become the beta=1 epilogue of the input-gradient GEMMs (addmm),
#   * the bias gradients of dense / dense_2 / dense_1 come out of pr_add_ln_bwd_bias_f32 / pr_act_bwd_bias_f32 as
#     per-CTA partial column sums instead of ATen sum(0) passes that re-read the whole gradient tensor,
#   * q/k/v weight gradients are produced by ONE GEMM into a [3D, D] buffer and returned as views.
def _raw_add_ln_fwd(h, res, gamma, beta, eps, p_pre, seed, stream_pre):
    rows, D = h.numel() // gamma.numel(), gamma.numel()
    y = torch.empty_like(h)
    mean = torch.empty(rows, device=h.device, dtype=torch.float32)
    rstd = torch.empty(rows, device=h.device, dtype=torch.float32)
    with _prof("add_ln_fwd", h):
        _lib.check(_L().pr_add_ln_fwd_f32(_p(h), 0, rows, _p(res), 0, _p(gamma), _p(beta), eps, rows, D, p_pre, 0.0, seed,
                                          stream_pre, 0, _p(y), _p(mean), _p(rstd), _stream(h)), "pr_add_ln_fwd_f32")
    _count()
    return y, mean, rstd


def _raw_add_ln_bwd_bias(dy, h, res, gamma, mean, rstd, p_pre, seed, stream_pre):
    """-> dh (dropout applied), dres (== dh when p_pre == 0), dgamma, dbeta, dbias (column sums of dh)"""
    rows, D = h.numel() // gamma.numel(), gamma.numel()
    n_part = _L().pr_add_ln_bwd_partials(rows, D)
    partials = torch.empty(3, n_part, D, device=dy.device, dtype=torch.float32)
    dh = torch.empty_like(h)
    dres = torch.empty_like(h) if p_pre > 0.0 else None
    with _prof("add_ln_bwd", dy):
        _lib.check(_L().pr_add_ln_bwd_bias_f32(_p(dy), _p(h), 0, rows, _p(res), 0, _p(gamma), _p(mean), _p(rstd), rows, D,
                                               p_pre, 0.0, seed, stream_pre, 0, _p(dh), 0, 0, _p(dres), _p(partials), n_part,
                                               _stream(dy)), "pr_add_ln_bwd_bias_f32")
    out = torch.empty(3, D, device=dy.device, dtype=torch.float32)
    with _prof("colsum", dy):
        _lib.check(_L().pr_colsum_f32(_p(partials), 3, n_part, D, _p(out), _stream(dy)), "pr_colsum_f32")
    _count(2)
    return dh, (dres if dres is not None else dh), out[0], out[1], out[2]


def _raw_ln_z_bwd_bias(dy, z, gamma, mean, rstd, p_pre, seed, stream_pre):
    """backward of LayerNorm(z), z = drop(h) + res written by gemm_drop_add -> dh (dropout applied), dz (gradient of the
    residual branch; == dh when p_pre == 0), dgamma, dbeta, dbias (column sums of dh).  Reads dy and z only."""
    rows, D = z.numel() // gamma.numel(), gamma.numel()
    n_part = _L().pr_add_ln_bwd_partials(rows, D)
    partials = torch.empty(3, n_part, D, device=dy.device, dtype=torch.float32)
    dh = torch.empty_like(z)
    dz = torch.empty_like(z) if p_pre > 0.0 else None
    with _prof("add_ln_bwd", dy):
        _lib.check(_L().pr_add_ln_bwd_bias_z_f32(_p(dy), _p(z), _p(gamma), _p(mean), _p(rstd), rows, D, p_pre, seed, stream_pre,
                                                 _p(dh), _p(dz), _p(partials), n_part, _stream(dy)), "pr_add_ln_bwd_bias_z_f32")
    out = torch.empty(3, D, device=dy.device, dtype=torch.float32)
    with _prof("colsum", dy):
        _lib.check(_L().pr_colsum_f32(_p(partials), 3, n_part, D, _p(out), _stream(dy)), "pr_colsum_f32")
    _count(2)
    return dh, (dz if dz is not None else dh), out[0], out[1], out[2]


def _raw_act_bwd_bias(x, dy, act):
    rows, cols = x.numel() // x.shape[-1], x.shape[-1]
    n_part = _L().pr_act_bwd_bias_partials(rows, cols)
    partials = torch.empty(n_part, cols, device=x.device, dtype=torch.float32)
    dx = torch.empty_like(x)
    with _prof("act_bwd", x):
        _lib.check(_L().pr_act_bwd_bias_f32(_p(x), _p(dy), rows, cols, act, _p(dx), _p(partials), n_part, _stream(x)),
                   "pr_act_bwd_bias_f32")
    db = torch.empty(1, cols, device=x.device, dtype=torch.float32)
    with _prof("colsum", x):
        _lib.check(_L().pr_colsum_f32(_p(partials), 1, n_part, cols, _p(db), _stream(x)), "pr_colsum_f32")
    _count(2)
    return dx, db[0]


def colsum_rows(x2):
    """[M, C] -> [C] column sums (bias gradient), one pass over x2 at HBM speed, deterministic (pr_colsum_rows_f32)"""
    _req(x2, torch.float32, "x2")
    M, C = x2.shape
    n_part = _L().pr_colsum_rows_partials(M, C)
    partials = torch.empty(max(n_part, 1), C, device=x2.device, dtype=torch.float32)
    out = torch.empty(C, device=x2.device, dtype=torch.float32)
    with _prof("colsum_rows", x2):
        _lib.check(_L().pr_colsum_rows_f32(_p(x2), M, C, _p(partials), n_part, _p(out), _stream(x2)), "pr_colsum_rows_f32")
    _count(2)
    return out


WGRAD_SPLIT = int(os.environ.get("PR_WGRAD_SPLIT", "8"))
# Linear layers of the encoder: "tc" = our CTA-pair tcgen05 GEMM (pr_gemm_tf32) for forward, input-gradient and weight-gradient
# GEMMs whenever TF32 matmuls are allowed (torch.backends.cuda.matmul.allow_tf32, the reference's torch-1.10 default);
# "cublas" = torch.addmm / mm (kept for A/B runs and for strict-fp32 parity runs, where allow_tf32 is off).
LINEAR_IMPL = os.environ.get("PR_LINEAR", "tc").lower()


FUSE_ACT_BWD = os.environ.get("PR_FUSE_ACT_BWD", "1") == "1"
# dense / dense_2 write z = dropout(x W^T + b) + residual from the GEMM epilogue (pr_gemm_tf32_drop); LayerNorm forward then reads one
# tensor instead of two and its backward two instead of three (PR_FUSE_LN_Z=0: the separate-kernel form, for A/B runs)
FUSE_LN_Z = os.environ.get("PR_FUSE_LN_Z", "1") == "1"


def _use_tc(*dims):
    return LINEAR_IMPL == "tc" and bool(torch.backends.cuda.matmul.allow_tf32) and all(d % 4 == 0 for d in dims)


def _linear_fwd(x2, w, b):
    """x2 [M, K] @ w[N, K]^T + b   (nn.Linear, layers.py:586-588, 613, 669)"""
    if _use_tc(x2.shape[1], w.shape[0]):
        return gemm(x2, w, bias=b, debias=True)
    return torch.addmm(b, x2, w.t())


def _linear_dgrad(dy2, w, add=None):
    """dy2 [M, N] @ w[N, K] (+ add): the input gradient of nn.Linear; `add` folds a residual gradient in"""
    if _use_tc(w.shape[0], w.shape[1]):
        if add is None:
            return gemm(dy2, w, b_mn=True, debias=True)
        return gemm(dy2, w, b_mn=True, aux=add, epi=GEMM_ADD, debias=True)
    return dy2.mm(w) if add is None else torch.addmm(add, dy2, w)


def _wgrad_splits(n_out, n_in, rows):
    tiles = ((n_out + 255) // 256) * ((n_in + 255) // 256)
    kb = (rows + 31) // 32
    s = max(1, min(74 // max(tiles, 1), kb // 8))           # ~one work item per CTA pair, >= 8 k-blocks each
    per = (kb + s - 1) // s
    return (kb + per - 1) // per                              # no empty split


def _wgrad(dy2, x2):
    """dW = dy2^T @ x2 for dy2 [M, out], x2 [M, in] (M = B*L is the contraction).  The output is a few hundred tiles at
    most, so one cuBLAS GEMM leaves most SMs idle or falls back to a slow split; an explicit split-K (batched GEMM over
    WGRAD_SPLIT slabs of M + one fixed-order sum) measured 1.3-2.0x faster at M = 81920 (profiles/r01h_rowkernels_ab.md)
    and stays deterministic."""
    M = dy2.shape[0]
    if _use_tc(dy2.shape[1], x2.shape[1]):
        return gemm(dy2, x2, a_mn=True, b_mn=True, splits=_wgrad_splits(dy2.shape[1], x2.shape[1], M), debias=True)
    S = WGRAD_SPLIT
    if S > 1 and M % S == 0 and M // S >= 2048:
        return torch.bmm(dy2.view(S, M // S, -1).transpose(1, 2), x2.view(S, M // S, -1)).sum(0)
    return dy2.t().mm(x2)


class LinearFn(torch.autograd.Function):
    """y = act(x W^T + b) for a stand-alone nn.Linear (the CLIP ViT item encoder's projections, REC/model/load.py:90-117): forward,
    input gradient and weight gradient on pr_gemm_tf32; bias + activation fused into the forward epilogue."""

    @staticmethod
    def forward(ctx, x, w, b, act):
        K = w.shape[1]
        x2 = x.reshape(-1, K)
        need = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        pre = None
        if act is None:
            y = gemm(x2, w, bias=b, debias=True)
        elif need:
            y, pre = gemm(x2, w, bias=b, epi=GEMM_ACT, act=act, want_pre=True, debias=True)
        else:
            y = gemm(x2, w, bias=b, epi=GEMM_ACT, act=act, debias=True)
        ctx.save_for_backward(x2, w, pre)
        ctx.act, ctx.xshape, ctx.has_bias = act, x.shape, b is not None
        return y.view(*x.shape[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, w, pre = ctx.saved_tensors
        dy2 = dy.reshape(-1, w.shape[0]).contiguous()
        if ctx.act is not None:
            d = torch.empty_like(dy2)
            with _prof("act_bwd", dy2):
                _lib.check(_L().pr_act_bwd_f32(_p(pre), _p(dy2), dy2.numel(), ctx.act, _p(d), _stream(dy2)), "pr_act_bwd_f32")
            _count()
            dy2 = d
        dx = gemm(dy2, w, b_mn=True, debias=True).view(ctx.xshape) if ctx.needs_input_grad[0] else None
        dw = None
        if ctx.needs_input_grad[1]:
            dw = gemm(dy2, x2, a_mn=True, b_mn=True, splits=_wgrad_splits(w.shape[0], w.shape[1], x2.shape[0]), debias=True)
        db = (colsum_rows(dy2) if dy2.shape[1] % 4 == 0 else dy2.sum(0)) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        return dx, dw, db, None


def linear(x, weight, bias=None, act=None):
    """nn.Linear [+ activation] on our tcgen05 GEMM when TF32 matmuls are allowed (torch.backends.cuda.matmul.allow_tf32), else
    the strict-fp32 library path (torch F.linear + our activation kernel) used by the 1e-4 parity runs."""
    act_id = None if act is None else ACT_IDS[act]
    if _use_tc(weight.shape[0], weight.shape[1]) and x.is_cuda:
        return LinearFn.apply(x.contiguous(), weight, bias, act_id)
    y = torch.nn.functional.linear(x, weight, bias)
    return y if act is None else activation(y.contiguous(), act)


class TransformerLayerFn(torch.autograd.Function):
    """x [B,L,D] -> FeedForward(MultiHeadAttention(x))  (layers.py:700-703), one autograd node."""

    @staticmethod
    def forward(ctx, x, key_ids, wq, bq, wk, bk, wv, bv, wo, bo, g1, be1, w1, b1, w2, b2, g2, be2, n_heads, causal, eps,
                p_attn, p_hid, act, seed, site):
        _req(x, torch.float32, "x")
        B, L, D = x.shape
        x2 = x.view(B * L, D)
        wqkv = torch.cat([wq, wk, wv], 0)
        bqkv = torch.cat([bq, bk, bv], 0)
        qkv = _linear_fwd(x2, wqkv, bqkv).view(B, L, 3 * D)                          # layers.py:586-588
        tf32 = bool(torch.backends.cuda.matmul.allow_tf32) and L <= 32 and (D // n_heads) % 32 == 0 and \
            ((D // n_heads) <= 128 or (D // n_heads) % 128 == 0)
        ctxt = torch.empty(B, L, D, device=x.device, dtype=torch.float32)
        probs = torch.empty(B, n_heads, L, L, device=x.device, dtype=torch.float32)
        base = qkv.data_ptr()
        fwd = _L().pr_sasrec_attn_fwd_tf32 if tf32 else _L().pr_sasrec_attn_fwd_f32
        with _prof("attn_fwd", qkv):
            _lib.check(fwd(base, base + 4 * D, base + 8 * D, 3 * D, _p(key_ids), B, L, n_heads, D // n_heads, int(causal), p_attn,
                           seed, site, _p(ctxt), _p(probs), _stream(qkv)), "pr_sasrec_attn_fwd")
        _count()
        zmode = FUSE_LN_Z and _use_tc(D, D, w1.shape[0])     # dense / dense_2 write z = dropout(linear) + residual themselves
        if zmode:
            h = gemm_drop_add(ctxt.view(B * L, D), wo, bo, x2, p_hid, seed, site + 1)      # :613-614 and the residual add of :615
            a, mean1, rstd1 = _raw_add_ln_fwd(h, None, g1, be1, eps, 0.0, seed, site + 1)  # :615
        else:
            h = _linear_fwd(ctxt.view(B * L, D), wo, bo)                                   # :613
            a, mean1, rstd1 = _raw_add_ln_fwd(h, x2, g1, be1, eps, p_hid, seed, site + 1)  # :614-615
        if _use_tc(a.shape[1], w1.shape[0]):                                          # dense_1 + activation in ONE kernel (h1 and act(h1) both kept)
            gl, h1 = gemm(a, w1, bias=b1, epi=GEMM_ACT, act=act, want_pre=True, debias=True)   # :666-667
        else:
            h1 = torch.addmm(b1, a, w1.t())                                            # :666
            gl = torch.empty_like(h1)
            with _prof("act_fwd", h1):
                _lib.check(_L().pr_act_fwd_f32(_p(h1), h1.numel(), act, _p(gl), _stream(h1)), "pr_act_fwd_f32")
            _count()
        if zmode:
            h2 = gemm_drop_add(gl, w2, b2, a, p_hid, seed, site + 2)                       # :669-670 and the residual add of :671
            y, mean2, rstd2 = _raw_add_ln_fwd(h2, None, g2, be2, eps, 0.0, seed, site + 2)
        else:
            h2 = _linear_fwd(gl, w2, b2)                                                   # :669
            y, mean2, rstd2 = _raw_add_ln_fwd(h2, a, g2, be2, eps, p_hid, seed, site + 2)  # :670-671
        ctx.save_for_backward(x, qkv, probs, ctxt, h, a, mean1, rstd1, h1, gl, h2, mean2, rstd2, wqkv, wo, w1, w2, g1, g2)
        ctx.cfg = (B, L, D, n_heads, int(causal), eps, p_attn, p_hid, act, seed, site, tf32, zmode)
        return y.view(B, L, D)

    @staticmethod
    def backward(ctx, dy):
        x, qkv, probs, ctxt, h, a, mean1, rstd1, h1, gl, h2, mean2, rstd2, wqkv, wo, w1, w2, g1, g2 = ctx.saved_tensors
        B, L, D, n_heads, causal, eps, p_attn, p_hid, act, seed, site, tf32, zmode = ctx.cfg
        M = B * L
        dy = dy.contiguous().view(M, D)
        x2 = x.view(M, D)
        # ---- feed-forward block
        if zmode:                                                          # h2 / h hold z: two tensors read instead of three
            dh2, da_res, dg2, dbe2, db2 = _raw_ln_z_bwd_bias(dy, h2, g2, mean2, rstd2, p_hid, seed, site + 2)
        else:
            dh2, da_res, dg2, dbe2, db2 = _raw_add_ln_bwd_bias(dy, h2, a, g2, mean2, rstd2, p_hid, seed, site + 2)
        dw2 = _wgrad(dh2, gl)
        if _use_tc(w2.shape[0], w2.shape[1]) and FUSE_ACT_BWD:
            # input gradient of dense_2, the activation's backward and the bias gradient of dense_1 in one kernel
            dh1, db1 = gemm(dh2, w2, b_mn=True, aux=h1, epi=GEMM_ACT_BWD, act=act, want_colsum=True, debias=True)
        else:
            dgl = _linear_dgrad(dh2, w2)
            dh1, db1 = _raw_act_bwd_bias(h1, dgl, act)
        dw1 = _wgrad(dh1, a)
        da = _linear_dgrad(dh1, w1, add=da_res)                            # residual grad folded into the GEMM epilogue
        # ---- attention block
        if zmode:
            dh, dx_res, dg1, dbe1, dbo = _raw_ln_z_bwd_bias(da, h, g1, mean1, rstd1, p_hid, seed, site + 1)
        else:
            dh, dx_res, dg1, dbe1, dbo = _raw_add_ln_bwd_bias(da, h, x2, g1, mean1, rstd1, p_hid, seed, site + 1)
        dwo = _wgrad(dh, ctxt.view(M, D))
        dctx = _linear_dgrad(dh, wo)
        dqkv = torch.empty_like(qkv)
        base, gbase = qkv.data_ptr(), dqkv.data_ptr()
        bwd = _L().pr_sasrec_attn_bwd_tf32 if tf32 else _L().pr_sasrec_attn_bwd_f32
        with _prof("attn_bwd", qkv):
            _lib.check(bwd(base, base + 4 * D, base + 8 * D, 3 * D, _p(probs), _p(dctx), B, L, n_heads, D // n_heads, causal,
                           p_attn, seed, site, gbase, gbase + 4 * D, gbase + 8 * D, 3 * D, _stream(qkv)), "pr_sasrec_attn_bwd")
        _count()
        dqkv2 = dqkv.view(M, 3 * D)
        dwqkv = _wgrad(dqkv2, x2)                                           # one GEMM for the three projections
        dbqkv = colsum_rows(dqkv2) if dqkv2.shape[1] % 4 == 0 else dqkv2.sum(0)
        dx = _linear_dgrad(dqkv2, wqkv, add=dx_res).view(B, L, D)          # residual grad folded into the GEMM epilogue
        return (dx, None, dwqkv[:D], dbqkv[:D], dwqkv[D:2 * D], dbqkv[D:2 * D], dwqkv[2 * D:], dbqkv[2 * D:], dwo, dbo, dg1, dbe1,
                dw1, db1, dw2, db2, dg2, dbe2, None, None, None, None, None, None, None, None)


# ------------------------------------------------------------------------------------------- K9 score + top-k
def score_topk(seq_out, item_feature, k, hist_u=None, hist_i=None, mask_col0=True):
    """Fused  (seq_out @ item_feature.T) -> mask(col 0, history) -> top-k  on tcgen05 tensor cores.
    Returns (values [B_e,k] fp32 descending, indices [B_e,k] int64); the [B_e,N] scores are never materialised."""
    _req(seq_out, torch.float32, "seq_out")
    _req(item_feature, torch.float32, "item_feature")
    B_e, D = seq_out.shape
    N = item_feature.shape[0]
    n_hist = 0
    if hist_u is not None and hist_u.numel():
        _req(hist_u, torch.int64, "hist_u"); _req(hist_i, torch.int64, "hist_i")
        n_hist = hist_u.numel()
    ws_bytes = _L().pr_score_topk_workspace_bytes(B_e, N, k)
    if ws_bytes == 0:
        raise _lib.PixelRecB200Error(f"score_topk: unsupported shape B_e={B_e} N={N} k={k}")
    ws = torch.empty(ws_bytes, device=seq_out.device, dtype=torch.uint8)
    val = torch.empty(B_e, k, device=seq_out.device, dtype=torch.float32)
    idx = torch.empty(B_e, k, device=seq_out.device, dtype=torch.int64)
    with _prof("score_topk", seq_out):
        _lib.check(_L().pr_score_topk_f32(_p(seq_out), B_e, _p(item_feature), N, D, _p(hist_u) if n_hist else None,
                                          _p(hist_i) if n_hist else None, n_hist, int(bool(mask_col0)), int(k), _p(val),
                                          _p(idx), _p(ws), ws_bytes, _stream(seq_out)), "pr_score_topk_f32")
    _count(4 if n_hist else 3)
    return val, idx


def table_norm_max(item_feature):
    """device scalar max_j |item_feature[j]| -- the bound score_topk_exact needs; compute once per evaluation"""
    _req(item_feature, torch.float32, "item_feature")
    out = torch.empty(1, device=item_feature.device, dtype=torch.float32)
    with _prof("table_norm_max", item_feature):
        _lib.check(_L().pr_table_norm_max_f32(_p(item_feature), item_feature.shape[0], item_feature.shape[1], _p(out),
                                              _stream(item_feature)), "pr_table_norm_max_f32")
    _count()
    return out


def score_topk_exact(seq_out, item_feature, k, hist_u=None, hist_i=None, mask_col0=True, w_norm_max=None):
    """score_topk whose ids are those of an fp32 ranking (collector.py:133): TF32 tensor-core pass for 32 candidates per row,
    fp32 re-score + rank, whole-catalog fp32 fallback for rows that cannot be proven complete.  k <= 16, N >= 32.
    Returns (values, indices, n_fallback_rows [1] int32 on the device)."""
    _req(seq_out, torch.float32, "seq_out")
    _req(item_feature, torch.float32, "item_feature")
    B_e, D = seq_out.shape
    N = item_feature.shape[0]
    n_hist = 0
    if hist_u is not None and hist_u.numel():
        _req(hist_u, torch.int64, "hist_u"); _req(hist_i, torch.int64, "hist_i")
        n_hist = hist_u.numel()
    ws_bytes = _L().pr_score_topk_exact_workspace_bytes(B_e, N, k)
    if ws_bytes == 0 or N < 32:
        raise _lib.PixelRecB200Error(f"score_topk_exact: unsupported shape B_e={B_e} N={N} k={k} (k <= 16, N >= 32)")
    ws = torch.empty(ws_bytes, device=seq_out.device, dtype=torch.uint8)
    val = torch.empty(B_e, k, device=seq_out.device, dtype=torch.float32)
    idx = torch.empty(B_e, k, device=seq_out.device, dtype=torch.int64)
    nfb = torch.zeros(1, device=seq_out.device, dtype=torch.int32)
    with _prof("score_topk_exact", seq_out):
        _lib.check(_L().pr_score_topk_exact_f32(_p(seq_out), B_e, _p(item_feature), N, D, _p(hist_u) if n_hist else None,
                                                _p(hist_i) if n_hist else None, n_hist, int(bool(mask_col0)), int(k),
                                                _p(w_norm_max), _p(val), _p(idx), _p(nfb), _p(ws), ws_bytes, _stream(seq_out)),
                   "pr_score_topk_exact_f32")
    _count((4 if n_hist else 3) + 2 + (0 if w_norm_max is not None else 1))
    return val, idx, nfb


GEMM_STORE, GEMM_ADD, GEMM_ACT, GEMM_ACT_BWD = 0, 1, 2, 3


def gemm(A, B, a_mn=False, b_mn=False, bias=None, aux=None, epi=GEMM_STORE, act=None, out=None, want_pre=False, splits=1,
         want_colsum=False, debias=False):
    """out[M, N] = epilogue(A . B^T) on the CTA-pair tcgen05 GEMM (pr_gemm_tf32, csrc/gemm.cu) -- forward, input-gradient
    and weight-gradient GEMMs of nn.Linear (REC/model/layers.py:586-588, 613, 666, 669) without a transposed copy:
      A: [M, K] (a_mn=False) or [K, M] (a_mn=True);  B: [N, K] (b_mn=False, nn.Linear weight layout) or [K, N] (b_mn=True).
    epi / act / aux / want_pre / want_colsum: see include/pixelrec_b200.h.  splits > 1: split-K, reduced here in fixed order.
    debias: compensate the systematic shrink of TF32 operand truncation (PR_GEMM_DEBIAS; the layer path turns it on).
    Returns out, or (out, pre) with want_pre, or (out, colsum [N]) with want_colsum."""
    _req(A, torch.float32, "A")
    _req(B, torch.float32, "B")
    if A.dim() != 2 or B.dim() != 2:
        raise ValueError("gemm: A and B must be 2-D")
    K, M = (A.shape if a_mn else (A.shape[1], A.shape[0]))
    Kb, N = (B.shape if b_mn else (B.shape[1], B.shape[0]))
    if K != Kb:
        raise ValueError(f"gemm: contraction lengths differ (A {tuple(A.shape)}, B {tuple(B.shape)})")
    dev = A.device
    if out is None:
        out = torch.empty(M, N, device=dev, dtype=torch.float32)
    else:
        _req(out, torch.float32, "out")
    pre = torch.empty(M, N, device=dev, dtype=torch.float32) if want_pre else None
    if aux is not None:
        _req(aux, torch.float32, "aux")
    act_id = -1 if act is None else (act if isinstance(act, int) else ACT_IDS[act])
    L_ = _L()
    partials = None
    if want_colsum:
        n_part = L_.pr_gemm_colsum_rows(M)
        partials = torch.empty(n_part, N, device=dev, dtype=torch.float32)
    dst = out
    if splits > 1:
        dst = torch.empty(splits, M, N, device=dev, dtype=torch.float32)
    with _prof("gemm", A):
        _lib.check(L_.pr_gemm_tf32(_p(A), int(a_mn), A.stride(0), _p(B), int(b_mn), B.stride(0), M, N, K, _p(bias), _p(aux),
                                   int(epi), act_id, _p(dst), _p(pre), int(splits), _p(partials), int(bool(debias)), _stream(A)),
                   "pr_gemm_tf32")
    _count()
    if splits > 1:
        with _prof("gemm_splitk_reduce", A):
            _lib.check(L_.pr_gemm_splitk_reduce_f32(_p(dst), int(splits), M * N, _p(out), _stream(A)),
                       "pr_gemm_splitk_reduce_f32")
        _count()
    if want_colsum:
        cs = torch.empty(1, N, device=dev, dtype=torch.float32)
        with _prof("colsum", A):
            _lib.check(L_.pr_colsum_f32(_p(partials), 1, partials.shape[0], N, _p(cs), _stream(A)), "pr_colsum_f32")
        _count()
        return out, cs[0]
    return (out, pre) if want_pre else out


def gemm_drop_add(x2, w, bias, res, p_drop, seed, rng_stream, debias=True):
    """z = dropout(x2 @ w^T + bias) + res in ONE kernel (pr_gemm_tf32_drop: REC/model/layers.py:613-614 / 669-670 up to the
    LayerNorm): the keep bits are those pr_add_ln_fwd_f32 would draw for (seed, rng_stream), so the LayerNorm that follows
    reads one tensor instead of two and its backward (pr_add_ln_bwd_bias_z_f32) regenerates the same mask."""
    _req(x2, torch.float32, "x2"); _req(w, torch.float32, "w"); _req(res, torch.float32, "res")
    M, K = x2.shape
    N = w.shape[0]
    if w.shape[1] != K or res.shape != (M, N):
        raise ValueError("gemm_drop_add: shapes do not match")
    out = torch.empty(M, N, device=x2.device, dtype=torch.float32)
    with _prof("gemm", x2):
        _lib.check(_L().pr_gemm_tf32_drop(_p(x2), 0, x2.stride(0), _p(w), 0, w.stride(0), M, N, K, _p(bias), _p(res), _p(out),
                                          int(bool(debias)), float(p_drop), int(seed), int(rng_stream), _stream(x2)),
                   "pr_gemm_tf32_drop")
    _count()
    return out


def score_prepare_f16(x, status=None):
    """fp16 copy (round to nearest even, saturating; status bit 2 = a value was beyond +-65504) of an fp32 CUDA tensor:
    the operand format of score_topk_f16.  Convert the item table once per evaluation, not per batch."""
    _req(x, torch.float32, "x")
    if x.numel() % 4:
        raise ValueError("score_prepare_f16: element count must be a multiple of 4")
    out = torch.empty(x.shape, device=x.device, dtype=torch.float16)
    with _prof("score_prepare_f16", x):
        _lib.check(_L().pr_score_prepare_f16(_p(x), x.numel(), _p(out), _p(status), _stream(x)), "pr_score_prepare_f16")
    _count()
    return out


def score_topk_f16(seq_out, item_feature_f16, k, hist_u=None, hist_i=None, mask_col0=True, status=None):
    """score_topk with fp16 operands (kind::f16 MMAs: the mantissa width of TF32 at twice the rate and half the operand
    bytes).  seq_out stays fp32 (converted inside); item_feature_f16 comes from score_prepare_f16.  D % 64 == 0."""
    _req(seq_out, torch.float32, "seq_out")
    _req(item_feature_f16, torch.float16, "item_feature_f16")
    B_e, D = seq_out.shape
    N = item_feature_f16.shape[0]
    n_hist = 0
    if hist_u is not None and hist_u.numel():
        _req(hist_u, torch.int64, "hist_u"); _req(hist_i, torch.int64, "hist_i")
        n_hist = hist_u.numel()
    ws_bytes = _L().pr_score_topk_f16_workspace_bytes(B_e, N, D, k)
    if ws_bytes == 0:
        raise _lib.PixelRecB200Error(f"score_topk_f16: unsupported shape B_e={B_e} N={N} D={D} k={k}")
    ws = torch.empty(ws_bytes, device=seq_out.device, dtype=torch.uint8)
    val = torch.empty(B_e, k, device=seq_out.device, dtype=torch.float32)
    idx = torch.empty(B_e, k, device=seq_out.device, dtype=torch.int64)
    with _prof("score_topk_f16", seq_out):
        _lib.check(_L().pr_score_topk_f16(_p(seq_out), B_e, _p(item_feature_f16), N, D, _p(hist_u) if n_hist else None,
                                          _p(hist_i) if n_hist else None, n_hist, int(bool(mask_col0)), int(k), _p(val),
                                          _p(idx), _p(ws), ws_bytes, _p(status), _stream(seq_out)), "pr_score_topk_f16")
    _count(5 if n_hist else 4)
    return val, idx


def score_ce(seq_out, item_feature, target, mask_col0=True):
    """Full-catalog softmax cross-entropy on the tcgen05 scoring pipeline (extension, forward only; include/pixelrec_b200.h):
    returns (lse [B_e], target logit [B_e], nll [B_e]) without materialising the [B_e, N] logits."""
    _req(seq_out, torch.float32, "seq_out")
    _req(item_feature, torch.float32, "item_feature")
    _req(target, torch.int64, "target")
    B_e, D = seq_out.shape
    N = item_feature.shape[0]
    ws_bytes = _L().pr_score_ce_workspace_bytes(B_e, N)
    if ws_bytes == 0:
        raise _lib.PixelRecB200Error(f"score_ce: unsupported shape B_e={B_e} N={N}")
    ws = torch.empty(ws_bytes, device=seq_out.device, dtype=torch.uint8)
    out = torch.empty(3, B_e, device=seq_out.device, dtype=torch.float32)
    with _prof("score_ce", seq_out):
        _lib.check(_L().pr_score_ce_f32(_p(seq_out), B_e, _p(item_feature), N, D, _p(target), int(bool(mask_col0)), _p(out[0]),
                                        _p(out[1]), _p(out[2]), _p(ws), ws_bytes, _stream(seq_out)), "pr_score_ce_f32")
    _count(3)
    return out[0], out[1], out[2]


CE_CHUNK = int(os.environ.get("PR_CE_CHUNK", "8192"))      # catalog columns per backward chunk (multiple of 4)


class ScoreCEFn(torch.autograd.Function):
    """nll[r] = logsumexp_c <seq_out[r], W[c]> - <seq_out[r], W[target[r]]> over the whole catalog (padding item excluded), forward
    AND backward without ever holding the [B_e, N] logits (extension: the reference trains with sampled negatives, sasrec.py:88-92;
    oracle = F.cross_entropy, oracle/sasrec_np.py full_catalog_ce).
      forward : pr_score_ce_f32 -- tcgen05 scoring GEMM with the online max / sum-of-exp in its epilogue;
      backward: the catalog in chunks of CE_CHUNK items: logits of the chunk recomputed by pr_gemm_tf32, turned in place into
                dS = dnll * (softmax - onehot) (pr_ce_grad_chunk_f32), then dX += dS W_c and dW_c = dS^T X on pr_gemm_tf32."""

    @staticmethod
    def forward(ctx, seq_out, item_feature, target, mask_col0):
        lse, tgt, nll = score_ce(seq_out, item_feature, target, mask_col0)
        ctx.save_for_backward(seq_out, item_feature, target, lse)
        ctx.mask_col0 = bool(mask_col0)
        return nll

    @staticmethod
    def backward(ctx, dnll):
        X, W, target, lse = ctx.saved_tensors
        B_e, D = X.shape
        N = W.shape[0]
        dnll = dnll.contiguous().float()
        need_x, need_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        dX = torch.zeros_like(X) if need_x else None
        dW = torch.empty_like(W) if need_w else None
        Xc, Wc = X.contiguous(), W.contiguous()
        C = max(4, CE_CHUNK // 4 * 4)
        for c0 in range(0, N, C):
            n = min(C, N - c0)
            n4 = (n + 3) // 4 * 4                       # the GEMM wants N % 4 == 0: pad columns read past the chunk
            Wch = Wc[c0:c0 + n]
            if n4 != n:                                 # ragged tail of the catalog: zero rows give exp(-lse) ~ 0 logits' worth of
                Wch = torch.cat([Wch, Wch.new_zeros(n4 - n, D)], 0)   # gradient; they are cut off below
            S = gemm(Xc, Wch)                           # [B_e, n4] logits of the chunk (TF32, like the forward)
            with _prof("ce_grad_chunk", S):
                _lib.check(_L().pr_ce_grad_chunk_f32(_p(S), S.stride(0), B_e, n4, c0, _p(lse), _p(target), _p(dnll),
                                                     int(ctx.mask_col0), _stream(S)), "pr_ce_grad_chunk_f32")
            _count()
            if n4 != n:
                S[:, n:] = 0.0
            if need_x:
                dX = gemm(S, Wch, b_mn=True, aux=dX, epi=GEMM_ADD)                  # dX += dS W_c   (contraction over the chunk)
            if need_w:
                if n4 == n:
                    gemm(S, Xc, a_mn=True, b_mn=True, out=dW[c0:c0 + n])              # dW_c = dS^T X  (contraction over the rows)
                else:
                    dW[c0:c0 + n] = gemm(S, Xc, a_mn=True, b_mn=True)[:n]
        return dX, dW, None, None


def score_ce_loss(seq_out, item_feature, target, mask_col0=True):
    """differentiable per-row full-catalog cross-entropy (see ScoreCEFn)"""
    return ScoreCEFn.apply(seq_out, item_feature, target, mask_col0)


# ------------------------------------------------------------------------------------------- peer-memory exchange
_CAI_TYPESTR = {torch.float32: "<f4", torch.int64: "<i8", torch.int32: "<i4", torch.uint8: "|u1"}


class _DevArray:
    """__cuda_array_interface__ view of raw device memory, so torch can alias it without owning it."""

    def __init__(self, ptr, shape, dtype, owner):
        self.owner = owner               # keeps the allocation alive as long as a tensor aliases it
        self.__cuda_array_interface__ = {"shape": tuple(int(x) for x in shape), "typestr": _CAI_TYPESTR[dtype],
                                         "data": (int(ptr), False), "version": 2, "strides": None}


class SharedBuffer:
    """Device memory that the other ranks of the box can map (pr_shared_alloc: cudaMalloc + CUDA IPC handle).
    `.ref` is its device address in THIS process, `.handle` the 64 bytes to ship to the peers."""

    def __init__(self, nbytes, device):
        import ctypes as C
        self.device = torch.device(device)
        self.nbytes = int(nbytes)
        _lib.check(_L().pr_set_device(self.device.index), "pr_set_device")
        _cur_dev[0] = self.device.index
        ptr = C.c_void_p()
        h = C.create_string_buffer(64)
        _lib.check(_L().pr_shared_alloc(self.nbytes, C.byref(ptr), h), "pr_shared_alloc")
        self.ref = int(ptr.value)
        self.handle = h.raw

    def tensor(self, shape, dtype):
        n = 1
        for x in shape:
            n *= int(x)
        if n * torch.empty((), dtype=dtype).element_size() > self.nbytes:
            raise ValueError("SharedBuffer.tensor: view larger than the allocation")
        return torch.as_tensor(_DevArray(self.ref, shape, dtype, self), device=self.device)

    def __del__(self):
        try:
            if getattr(self, "ref", None):
                _L().pr_shared_free(self.ref)
        except Exception:      # interpreter shutdown
            pass


def shared_open(handle, device):
    """Maps a peer's SharedBuffer (its 64-byte handle) into this process; returns the device address."""
    import ctypes as C
    device = torch.device(device)
    _lib.check(_L().pr_set_device(device.index), "pr_set_device")
    _cur_dev[0] = device.index
    ptr = C.c_void_p()
    _lib.check(_L().pr_shared_open(bytes(handle), C.byref(ptr)), "pr_shared_open")
    return int(ptr.value)


def gather_rows_peers(shard_table, G, N, D, idx, status=None):
    """out[r] = shard[idx[r] % G][idx[r] // G] read straight from the owners' memory (lookup + exchange in one kernel).
    shard_table: int64 CUDA tensor [G] of device addresses (own shard and the peers' mapped shards)."""
    _req(shard_table, torch.int64, "shard_table")
    _req(idx, torch.int64, "idx")
    R = idx.numel()
    out = torch.empty(*idx.shape, D, device=idx.device, dtype=torch.float32)
    with _prof("gather_rows_peers", idx):
        _lib.check(_L().pr_gather_rows_peers_f32(_p(shard_table), int(G), int(N), int(D), _p(idx), R, _p(out), _p(status),
                                                 _stream(idx)), "pr_gather_rows_peers_f32")
    _count()
    return out


def peer_barrier(flag_table, G, rank, epoch, status=None, epoch_dev=None):
    """flag_table: int64 CUDA tensor [G] of device addresses of every rank's uint64 flag array (pr_peer_barrier);
    epoch_dev: int64 [1] device counter used instead of `epoch` (CUDA-graph replay)"""
    _req(flag_table, torch.int64, "flag_table")
    with _prof("exchange_barrier", flag_table):
        _lib.check(_L().pr_peer_barrier(_p(flag_table), int(G), int(rank), int(epoch), _p(epoch_dev), _p(status),
                                        _stream(flag_table)), "pr_peer_barrier")
    _count()


def plan_inverse(plan, pad_slot):
    """request position -> slot in plan.uniq_ids (pad_slot for the positions the plan dropped); int64 [R]"""
    inv = torch.empty(max(plan.R, 1), device=plan.perm.device, dtype=torch.int64)
    with _prof("plan_inverse", plan.perm):
        _lib.check(_L().pr_plan_inverse(_p(plan.perm), _p(plan.seg_start), _p(plan.n_uniq), plan.R, int(pad_slot), _p(inv),
                                        _stream(plan.perm)), "pr_plan_inverse")
    _count()
    return inv


def gather_rows_peers_plan(shard_table, G, N, D, plan, pad_id, pad_slot, status=None):
    """[plan.max_uniq + 1, D]: row u = the owner's row of plan.uniq_ids[u] (u < n_uniq, read over NVLink), row pad_slot = pad row"""
    _req(shard_table, torch.int64, "shard_table")
    out = torch.empty(plan.max_uniq + 1, D, device=plan.perm.device, dtype=torch.float32)
    with _prof("gather_rows_peers", plan.perm):
        _lib.check(_L().pr_gather_rows_peers_plan_f32(_p(shard_table), int(G), int(N), int(D), _p(plan.uniq_ids), _p(plan.n_uniq),
                                                      plan.max_uniq, -1 if pad_id is None else int(pad_id), int(pad_slot), _p(out),
                                                      _p(status), _stream(plan.perm)), "pr_gather_rows_peers_plan_f32")
    _count()
    return out


def push_rows_peers_plan(rows, plan, G, rank, cap, recv_rows_table, recv_ids_table, counters, status=None):
    _req(rows, torch.float32, "rows")
    _req(counters, torch.int32, "counters")
    with _prof("push_rows_peers", rows):
        _lib.check(_L().pr_push_rows_peers_plan_f32(_p(rows), _p(plan.uniq_ids), _p(plan.n_uniq), plan.max_uniq, rows.shape[1],
                                                    int(G), int(rank), int(cap), _p(recv_rows_table), _p(recv_ids_table),
                                                    _p(counters), _p(status), _stream(rows)), "pr_push_rows_peers_plan_f32")
    _count()


def push_rows_peers(rows, ids, G, rank, cap, skip_id, recv_rows_table, recv_ids_table, counters, status=None):
    """rows[u] -> the owner's receive region of this rank (see include/pixelrec_b200.h); counters [G] int32 zero on entry."""
    _req(rows, torch.float32, "rows")
    _req(ids, torch.int64, "ids")
    _req(counters, torch.int32, "counters")
    U, D = rows.shape
    if ids.numel() != U:
        raise ValueError("push_rows_peers: one id per row")
    with _prof("push_rows_peers", rows):
        _lib.check(_L().pr_push_rows_peers_f32(_p(rows), _p(ids), U, D, int(G), int(rank), int(cap),
                                               -1 if skip_id is None else int(skip_id), _p(recv_rows_table),
                                               _p(recv_ids_table), _p(counters), _p(status), _stream(rows)),
                   "pr_push_rows_peers_f32")
    _count()


# ------------------------------------------------------------------------------------------- A1 on-device batches
def seq_batch_build(padded, sel, item_num, seed, status=None):
    """items [B,2,W], masked_index [B,W-1] for the windows `sel` of `padded` [n_seq, W] (trainset.py:52-75 on the GPU)."""
    _req(padded, torch.int64, "padded"); _req(sel, torch.int64, "sel")
    n_seq, W = padded.shape
    B = sel.numel()
    items = torch.empty(B, 2, W, device=padded.device, dtype=torch.int64)
    mask = torch.empty(B, W - 1, device=padded.device, dtype=torch.int64)
    with _prof("seq_batch_build", padded):
        _lib.check(_L().pr_seq_batch_build(_p(padded), n_seq, W, _p(sel), B, int(item_num), int(seed) & 0xFFFFFFFFFFFFFFFF,
                                           _p(items), _p(mask), _p(status), _stream(padded)), "pr_seq_batch_build")
    _count()
    return items, mask


def inv_sqrt(x):
    return 1.0 / math.sqrt(x)
