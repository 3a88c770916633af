"""CPU check of SIMT kernel LOGIC without a GPU: the device code of csrc/attn_long.cuh and csrc/peer.cuh is compiled for the
host on a thread-per-CUDA-thread emulation layer (tests/emu/emu_cuda.h: pthread barriers for __syncthreads / warp
collectives) and compared with the oracle.  This pins indexing, masking, reductions and the slot protocol; it says nothing
about performance or about PTX-level behaviour -- the GPU parity tests (tests/test_gpu_*.py) remain the gate."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest

from oracle import sasrec_np as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_P, _I, _LL = C.c_void_p, C.c_int, C.c_longlong


@pytest.fixture(scope="module")
def emu():
    out = os.path.join(tempfile.mkdtemp(prefix="pr_emu_"), "libemu.so")
    src = os.path.join(ROOT, "tests", "emu", "emu_kernels.cpp")
    r = subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-pthread", "-o", out, src], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lib = C.CDLL(out)
    lib.emu_attn_long_fwd.argtypes = [_P, _P, _P, _LL, _P, _I, _I, _I, _I, _I, _P, _P, _I]
    lib.emu_attn_long_bwd.argtypes = [_P, _P, _P, _LL, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _LL, _P, _I]
    lib.emu_attn_long_tc_fwd.argtypes = [_P, _P, _P, _LL, _P, _I, _I, _I, _I, _I, _P, _P, _I]
    lib.emu_attn_long_tc_bwd.argtypes = [_P, _P, _P, _LL, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _LL, _P, _I]
    lib.emu_gather_rows_peers.argtypes = [_P, _I, _LL, _I, _P, _LL, _P, _P, _I]
    lib.emu_push_rows_peers.argtypes = [_P, _P, _LL, _I, _I, _I, _LL, _LL, _P, _P, _P, _P, _I]
    return lib


@pytest.fixture(scope="module")
def emu_score():
    """csrc/score_kernels.cuh on the emulated Blackwell pipeline (tests/emu/emu_tc.h)"""
    out = os.path.join(tempfile.mkdtemp(prefix="pr_emu_"), "libemu_score.so")
    src = os.path.join(ROOT, "tests", "emu", "emu_score.cpp")
    r = subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-pthread", "-o", out, src], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lib = C.CDLL(out)
    lib.emu_score_topk_v2.argtypes = [_P, _LL, _P, _LL, _LL, _P, _P, _LL, _I, _I, _I, _I, _P, _P]
    lib.emu_score_topk_f16.argtypes = [_P, _LL, _P, _LL, _LL, _P, _P, _LL, _I, _I, _I, _I, _P, _P, _P, _I]
    lib.emu_f32_to_f16.argtypes = [_P, _LL, _P]
    lib.emu_score_ce_v2.argtypes = [_P, _LL, _P, _LL, _LL, _P, _I, _I, _I, _P, _P, _P]
    return lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _mask(key_ids, causal, B, L):
    valid = np.ones((B, L), bool) if key_ids is None else (key_ids != 0)
    tri = np.tril(np.ones((L, L), bool)) if causal else np.ones((L, L), bool)
    return np.where(valid[:, None, None, :] & tri[None, None], 0.0, -1e9)


@pytest.mark.parametrize("B,L,h,dh,causal,padded", [(1, 70, 2, 16, 0, False), (2, 197, 1, 8, 0, False), (1, 100, 2, 32, 1, True),
                                                  (1, 65, 1, 64, 0, True), (1, 33, 3, 4, 1, False), (1, 130, 1, 128, 0, False)])
def test_attn_long_fwd_bwd_emulated(emu, B, L, h, dh, causal, padded):
    g = np.random.default_rng(L * dh)
    D = h * dh
    qkv = g.standard_normal((B, L, 3 * D)).astype(np.float32)
    key_ids = None
    if padded:
        key_ids = g.integers(1, 50, size=(B, L)).astype(np.int64)
        key_ids[:, :L // 5] = 0                                   # left padding, as SEQTrainDataset produces
    q, k, v = qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:]
    mask = _mask(key_ids, causal, B, L)
    ref, cache = O.attn_core_fwd(q.astype(np.float64), k.astype(np.float64), v.astype(np.float64), mask, h)
    ctx = np.zeros((B, L, D), np.float32)
    lse = np.zeros((B * h, L), np.float32)
    kp = _ptr(key_ids) if key_ids is not None else None
    base = qkv.ctypes.data
    emu.emu_attn_long_fwd(base, base + 4 * D, base + 8 * D, 3 * D, kp, B, L, h, dh, causal, _ptr(ctx), _ptr(lse), 2)
    # query rows with no visible key (left padding under the causal mask) are don't-care: the reference's additive -1e9
    # makes them uniform in fp32 but not in the fp64 oracle; they never reach the loss
    live = (mask == 0).any(-1)[:, 0]                              # [B, L]
    assert live.any() and np.isfinite(ctx).all()
    assert np.abs(ctx - ref)[live].max() < 2e-5 * max(1.0, np.abs(ref).max())
    s = np.einsum("bihd,bjhd->bhij", q.reshape(B, L, h, dh).astype(np.float64), k.reshape(B, L, h, dh).astype(np.float64))
    s = s / np.sqrt(dh) + mask
    lse_ref = np.log(np.exp(s - s.max(-1, keepdims=True)).sum(-1)) + s.max(-1)
    lv = np.broadcast_to(live[:, None, :], (B, h, L))
    assert np.allclose(lse.reshape(B, h, L)[lv], lse_ref[lv], rtol=1e-5, atol=1e-4)
    dout = g.standard_normal((B, L, D)).astype(np.float32) * live[..., None]
    dq_r, dk_r, dv_r = O.attn_core_bwd(dout.astype(np.float64), cache)
    dqkv = np.zeros_like(qkv)
    delta = np.zeros((B * h, L), np.float32)
    gb = dqkv.ctypes.data
    emu.emu_attn_long_bwd(base, base + 4 * D, base + 8 * D, 3 * D, kp, _ptr(ctx), _ptr(lse), _ptr(dout), B, L, h, dh, causal,
                          gb, gb + 4 * D, gb + 8 * D, 3 * D, _ptr(delta), 2)
    for got, want, name in ((dqkv[..., :D], dq_r, "dq"), (dqkv[..., D:2 * D], dk_r, "dk"), (dqkv[..., 2 * D:], dv_r, "dv")):
        err = np.abs(got - want).max() / max(np.abs(want).max(), 1e-6)
        assert err < 5e-5, (name, err)


@pytest.mark.parametrize("B,L,h,dh,causal,padded", [(1, 70, 1, 32, 0, False), (1, 100, 2, 64, 1, True), (1, 197, 1, 64, 0, False)])
def test_attn_long_tc_fwd_emulated(emu, B, L, h, dh, causal, padded):
    """tensor-core forward (mma.sync TF32 emulated lane-exactly with the PTX fragment layouts): fragment index arithmetic,
    the accumulator->A-fragment key permutation, online softmax over 64-key blocks, masks, lse -- vs the fp64 oracle at TF32
    tolerance, and lse consistent with the fp32 kernels' (the backward consumes it)."""
    g = np.random.default_rng(L + dh)
    D = h * dh
    qkv = g.standard_normal((B, L, 3 * D)).astype(np.float32)
    key_ids = None
    if padded:
        key_ids = g.integers(1, 50, size=(B, L)).astype(np.int64)
        key_ids[:, :L // 5] = 0
    q, k, v = qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:]
    mask = _mask(key_ids, causal, B, L)
    ref, _ = O.attn_core_fwd(q.astype(np.float64), k.astype(np.float64), v.astype(np.float64), mask, h)
    live = (mask == 0).any(-1)[:, 0]
    ctx = np.zeros((B, L, D), np.float32)
    lse = np.zeros((B * h, L), np.float32)
    kp = _ptr(key_ids) if key_ids is not None else None
    base = qkv.ctypes.data
    assert emu.emu_attn_long_tc_fwd(base, base + 4 * D, base + 8 * D, 3 * D, kp, B, L, h, dh, causal, _ptr(ctx), _ptr(lse), 1) == 0
    assert np.isfinite(ctx).all()
    assert np.abs(ctx - ref)[live].max() < 3e-3 * max(1.0, np.abs(ref).max())          # TF32 operands
    ctx32 = np.zeros_like(ctx)
    lse32 = np.zeros_like(lse)
    emu.emu_attn_long_fwd(base, base + 4 * D, base + 8 * D, 3 * D, kp, B, L, h, dh, causal, _ptr(ctx32), _ptr(lse32), 1)
    lv = np.broadcast_to(live[:, None, :], (B, h, L))
    assert np.abs(lse - lse32).reshape(B, h, L)[lv].max() < 2e-2
    # tensor-core backward from the tensor-core forward's ctx / lse, vs the fp64 oracle
    _, cache = O.attn_core_fwd(q.astype(np.float64), k.astype(np.float64), v.astype(np.float64), mask, h)
    dout = g.standard_normal((B, L, D)).astype(np.float32) * live[..., None]
    dq_r, dk_r, dv_r = O.attn_core_bwd(dout.astype(np.float64), cache)
    dqkv = np.zeros_like(qkv)
    delta = np.zeros((B * h, L), np.float32)
    gb = dqkv.ctypes.data
    assert emu.emu_attn_long_tc_bwd(base, base + 4 * D, base + 8 * D, 3 * D, kp, _ptr(ctx), _ptr(lse), _ptr(dout), B, L, h, dh,
                                    causal, gb, gb + 4 * D, gb + 8 * D, 3 * D, _ptr(delta), 1) == 0
    for got, want, name in ((dqkv[..., :D], dq_r, "dq"), (dqkv[..., D:2 * D], dk_r, "dk"), (dqkv[..., 2 * D:], dv_r, "dv")):
        err = np.abs(got - want).max() / max(np.abs(want).max(), 1e-6)
        assert err < 5e-3, (name, err)


@pytest.mark.parametrize("G,N,D,R", [(1, 20, 8, 9), (2, 101, 16, 300), (4, 77, 36, 130), (8, 1003, 8, 2100)])
def test_gather_rows_peers_emulated(emu, G, N, D, R):
    g = np.random.default_rng(G * N)
    W = g.standard_normal((N, D)).astype(np.float32)
    shards = [np.ascontiguousarray(W[r::G]) if len(range(r, N, G)) else np.zeros((1, D), np.float32) for r in range(G)]
    table = (C.c_void_p * G)(*[s.ctypes.data for s in shards])
    idx = g.integers(0, N, size=R).astype(np.int64)
    idx[:2] = [0, N - 1]
    out = np.full((R, D), 7.0, np.float32)
    status = np.zeros(1, np.int32)
    emu.emu_gather_rows_peers(table, G, N, D, _ptr(idx), R, _ptr(out), _ptr(status), 3)
    assert np.array_equal(out, O.gather_rows(W, idx)) and status[0] == 0
    idx[1] = N + 3
    emu.emu_gather_rows_peers(table, G, N, D, _ptr(idx), R, _ptr(out), _ptr(status), 3)
    assert status[0] == 1 and (out[1] == 0).all() and np.array_equal(out[2:], W[idx[2:]])


@pytest.mark.parametrize("G,N,D,U,cap", [(2, 101, 8, 60, 60), (4, 1003, 16, 300, 300), (3, 50, 4, 40, 5)])
def test_push_rows_peers_emulated(emu, G, N, D, U, cap):
    """slot protocol of csrc/peer.cuh: every row lands exactly once in its owner's region of the sending rank (or is
    dropped and flagged when that region is full), ids are local rows, nothing is written outside the region"""
    g = np.random.default_rng(N + U)
    recv_rows = [np.full((G * cap, D), np.nan, np.float32) for _ in range(G)]
    recv_ids = [np.full(G * cap, -1, np.int64) for _ in range(G)]
    rt = (C.c_void_p * G)(*[a.ctypes.data for a in recv_rows])
    it = (C.c_void_p * G)(*[a.ctypes.data for a in recv_ids])
    status = np.zeros(1, np.int32)
    sent = {}
    for rank in range(G):
        ids = np.sort(g.choice(N, size=U, replace=False)).astype(np.int64)
        ids[0] = 0
        rows = g.standard_normal((U, D)).astype(np.float32)
        counters = np.zeros(G, np.int32)
        emu.emu_push_rows_peers(_ptr(rows), _ptr(ids), U, D, G, rank, cap, 0, rt, it, _ptr(counters), _ptr(status), 2)
        for o in range(G):
            assert counters[o] == ((ids[1:] % G) == o).sum()
        sent[rank] = (ids, rows, counters.copy())
    overflow = any((c > cap).any() for _, _, c in sent.values())
    assert bool(status[0] & 2) == overflow
    for o in range(G):
        for rank in range(G):
            ids, rows, counters = sent[rank]
            n = min(int(counters[o]), cap)
            reg_ids = recv_ids[o][rank * cap:(rank + 1) * cap]
            reg_rows = recv_rows[o][rank * cap:(rank + 1) * cap]
            assert (reg_ids[n:] == -1).all() and np.isnan(reg_rows[n:]).all()
            mine = {int(i): rows[u] for u, i in enumerate(ids) if i != 0 and i % G == o}
            got = set()
            for slot in range(n):
                gid = int(reg_ids[slot]) * G + o
                assert gid in mine and gid not in got and np.array_equal(reg_rows[slot], mine[gid])
                got.add(gid)
            assert len(got) == n and (n == len(mine) or overflow)


@pytest.mark.parametrize("B_e,N,D,k,splits,cluster", [(7, 40, 32, 32, 1, 1),          # fewer items than list slots, one k-block
                                                     (100, 300, 64, 10, 1, 1),       # one CTA, 2 tiles
                                                     (200, 1700, 160, 10, 2, 2),     # ring wraps 2.5x, TMEM buffers reused, cluster of 2
                                                     (130, 1000, 96, 20, 3, 2),      # K = 32 lists, ragged last tile / split
                                                     (500, 600, 64, 10, 2, 4)])      # cluster of 4 (B_e = 500: 4 m-tiles)
def test_score_topk_v2_pipeline_emulated(emu_score, B_e, N, D, k, splits, cluster):
    """score_topk2_kernel (branch-free 8-warp epilogue, optional table-tile multicast across a cluster) + merge + mask kernels
    on emulated mbarriers / asynchronous TMA / deferred MMAs / TMEM, against the oracle's full_sort_topk on integer operands
    (exact, massive ties).  A protocol error shows up as a wrong result, an over-arrival abort or a deadlock timeout."""
    g = np.random.default_rng(B_e + N)
    seq = g.integers(-3, 4, size=(B_e, D)).astype(np.float32)
    W = g.integers(-3, 4, size=(N, D)).astype(np.float32)
    hu = np.repeat(np.arange(B_e), 5).astype(np.int64)
    hi = g.integers(1, N, size=B_e * 5).astype(np.int64)
    scores = seq.astype(np.float64) @ W.astype(np.float64).T
    v_ref, i_ref = O.full_sort_topk(scores, hu, hi, k)
    val = np.zeros((B_e, k), np.float32)
    idx = np.zeros((B_e, k), np.int64)
    ns = emu_score.emu_score_topk_v2(_ptr(seq), B_e, _ptr(W), N, D, _ptr(hu), _ptr(hi), len(hu), 1, k, splits, cluster,
                                     _ptr(val), _ptr(idx))
    assert ns >= 1
    assert np.array_equal(idx, i_ref)
    assert np.array_equal(val.astype(np.float64), v_ref)


@pytest.mark.parametrize("B_e,N,D,splits,cluster", [(70, 300, 64, 1, 1), (200, 1500, 96, 3, 2)])
def test_score_ce_pipeline_emulated(emu_score, B_e, N, D, splits, cluster):
    """CE epilogue of the v2 scoring kernel (online max / sum of exp per thread, target-logit pick-up, partial merge) on the
    emulated pipeline vs the oracle's full_catalog_ce (extension: restates F.cross_entropy, not a reference path)."""
    g = np.random.default_rng(N + D)
    seq = g.standard_normal((B_e, D)).astype(np.float32)
    W = (0.3 * g.standard_normal((N, D))).astype(np.float32)
    target = g.integers(1, N, size=B_e).astype(np.int64)
    target[:3] = [1, N - 1, 255]
    lse_r, tl_r, nll_r = O.full_catalog_ce(seq.astype(np.float64) @ W.astype(np.float64).T, target)
    lse, tl, nll = (np.zeros(B_e, np.float32) for _ in range(3))
    ns = emu_score.emu_score_ce_v2(_ptr(seq), B_e, _ptr(W), N, D, _ptr(target), 1, splits, cluster, _ptr(lse), _ptr(tl), _ptr(nll))
    assert ns >= 1
    assert np.allclose(lse, lse_r, rtol=2e-5, atol=2e-5)
    assert np.allclose(tl, tl_r, rtol=2e-5, atol=2e-5)
    assert np.allclose(nll, nll_r, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("B_e,N,D,k,splits,cluster,ares", [(100, 300, 64, 10, 1, 1, 0), (200, 1700, 192, 10, 2, 2, 0),
                                                          (200, 1700, 192, 10, 2, 2, 1), (70, 1200, 512, 20, 1, 1, 1)])
def test_score_topk_f16_pipeline_emulated(emu_score, B_e, N, D, k, splits, cluster, ares):
    """fp16-operand mode of the v2 scoring kernel (kind::f16: 64 elements per 128-byte k-block, K = 16 per MMA) incl. the
    fp32 -> fp16 conversion kernel, on the emulated pipeline.  Small integers are exact in fp16, so results must equal the
    oracle bit for bit; a value beyond the fp16 range must raise the saturation flag."""
    g = np.random.default_rng(B_e * 3 + N)
    seq = g.integers(-3, 4, size=(B_e, D)).astype(np.float32)
    W = g.integers(-3, 4, size=(N, D)).astype(np.float32)
    hu = np.repeat(np.arange(B_e), 5).astype(np.int64)
    hi = g.integers(1, N, size=B_e * 5).astype(np.int64)
    v_ref, i_ref = O.full_sort_topk(seq.astype(np.float64) @ W.astype(np.float64).T, hu, hi, k)
    val = np.zeros((B_e, k), np.float32)
    idx = np.zeros((B_e, k), np.int64)
    status = np.zeros(1, np.int32)
    ns = emu_score.emu_score_topk_f16(_ptr(seq), B_e, _ptr(W), N, D, _ptr(hu), _ptr(hi), len(hu), 1, k, splits, cluster,
                                      _ptr(val), _ptr(idx), _ptr(status), ares)      # ares: seq_out tile resident, 3-stage table ring
    assert ns >= 1 and status[0] == 0
    assert np.array_equal(idx, i_ref)
    assert np.array_equal(val.astype(np.float64), v_ref)
    if cluster == 1:
        W2 = W.copy()
        W2[5, 3] = 1e6
        emu_score.emu_score_topk_f16(_ptr(seq), B_e, _ptr(W2), N, D, _ptr(hu), _ptr(hi), len(hu), 1, k, splits, cluster,
                                     _ptr(val), _ptr(idx), _ptr(status), ares)
        assert status[0] == 4


def test_f32_to_f16_routine_matches_numpy(emu_score):
    """the integer-only fp32 -> fp16 conversion used by pr_score_prepare_f16 (round to nearest even, subnormals, saturation
    instead of overflow to inf) against numpy's float16 cast on random bit patterns, scaled normals and edge values"""
    g = np.random.default_rng(0)
    parts = [g.standard_normal(200_000).astype(np.float32) * s for s in (1, 1e-3, 1e-5, 1e-7, 1e3, 3e4)]
    parts.append(g.integers(0, 2 ** 32, size=500_000, dtype=np.uint64).astype(np.uint32).view(np.float32))
    parts.append(np.array([0.0, -0.0, 65504, 65519.99, 65520, -65520, 1e10, 6.1035e-5, 6.0e-5, 5.96e-8, 2.98e-8, 2.9802322e-8,
                           2.99e-8, 1.0, 1.0009765625, 1.00048828125, 1.0014648, np.inf, -np.inf], np.float32))
    x = np.concatenate(parts)
    x = x[~np.isnan(x)]
    out = np.zeros(x.size, np.uint16)
    sat = emu_score.emu_f32_to_f16(_ptr(x), x.size, _ptr(out))
    with np.errstate(over="ignore"):
        h = x.astype(np.float16)
    ref = h.view(np.uint16).copy()
    over = np.isfinite(x) & np.isinf(h)
    ref[over] = (ref[over] & 0x8000) | 0x7BFF                      # saturate instead of inf
    assert sat == 1 and over.any()
    assert np.array_equal(out, ref)


@pytest.fixture(scope="module")
def emu_rows():
    """csrc/rows_ring.cuh on the emulated mbarrier / bulk-copy layer (tests/emu/emu_tc.h)"""
    out = os.path.join(tempfile.mkdtemp(prefix="pr_emu_"), "libemu_rows.so")
    src = os.path.join(ROOT, "tests", "emu", "emu_rows.cpp")
    r = subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-pthread", "-o", out, src], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lib = C.CDLL(out)
    lib.emu_scatter_add_rows_ring.argtypes = [_P, _I, _I, _P, _P, _P, _P, _LL, C.c_float, _P, _P, _I, _I, _I]
    lib.emu_scatter_plan.argtypes = [_P, _I, _LL, _LL, _P, _P, _P, _P, _P, _P, _I]
    return lib


def _np_plan(idx, N, pad):
    """what pr_scatter_plan emits: stable sort of (id, position) with padding / nothing dropped to the end"""
    key = np.where(idx == pad, N, idx)
    perm = np.argsort(key, kind="stable").astype(np.int32)
    sk = key[perm]
    nreal = int((sk < N).sum())
    starts = np.flatnonzero(np.r_[True, sk[1:nreal] != sk[:nreal - 1]]) if nreal else np.zeros(0, np.int64)
    uniq = sk[starts].astype(np.int32)
    seg = np.r_[starts, nreal].astype(np.int32)
    return perm, uniq, seg


@pytest.mark.parametrize("N,D,R,gr,grid,hot,nst,big", [(50, 64, 300, 4, 3, 0, 6, 0), (200, 512, 260, 32, 2, 0, 6, 0), (40, 384, 200, 8, 1, 0, 4, 0),
                                                       (30, 128, 500, 2, 7, 0.6, 6, 0), (300, 1024, 90, 32, 5, 0, 6, 0),
                                                       (64, 512, 170, 16, 2, 0.3, 4, 1), (20, 768, 40, 32, 3, 0, 6, 0),
                                                       (10, 512, 0, 32, 2, 0, 6, 0), (10, 512, 40, 32, 2, 1.0, 6, 0),
                                                       (100, 256, 333, 8, 4, 0.2, 3, 0),
                                                       # ring v2: look-ahead queue in shared memory (cp.async), 3- and 4-stage rings
                                                       (50, 64, 300, 4, 3, 0, 3, 2), (200, 512, 260, 32, 2, 0, 3, 2), (40, 384, 200, 8, 1, 0, 4, 2),
                                                       (30, 128, 500, 2, 7, 0.6, 3, 2), (300, 1024, 90, 32, 5, 0, 3, 2),
                                                       (10, 512, 0, 32, 2, 0, 3, 2), (10, 512, 40, 32, 2, 1.0, 3, 2),
                                                       (64, 512, 170, 16, 2, 0.3, 6, 2)])
def test_scatter_add_rows_ring_emulated(emu_rows, N, D, R, gr, grid, hot, nst, big):
    """TMA-staged segment reduce (rows_ring.cuh): ring protocol, group / run bookkeeping and the summation order, bit for bit
    against oracle.scatter_add_rows -- including one hot id that owns most rows, all-padding input and R = 0."""
    g = np.random.default_rng(N * 131 + D + R)
    idx = g.integers(0, N, size=R).astype(np.int64)
    if hot:
        idx[g.random(R) < hot] = 0 if hot == 1.0 else 7
    dO = g.standard_normal((max(R, 1), D)).astype(np.float32)[:R]
    perm, uniq, seg = _np_plan(idx, N, 0)
    U = len(uniq)
    max_uniq = max(1, min(R, N))
    uniq_b = np.full(max_uniq, -1, np.int32); uniq_b[:U] = uniq
    seg_b = np.full(max_uniq + 1, -1, np.int32); seg_b[:U + 1] = seg
    n_uniq = np.array([U], np.int32)
    rows = np.full((max_uniq, D), np.nan, np.float32)
    G = np.zeros((N, D), np.float32)
    dOc = np.ascontiguousarray(dO) if R else np.zeros((1, D), np.float32)
    assert emu_rows.emu_scatter_add_rows_ring(_ptr(dOc), D, gr, _ptr(perm), _ptr(uniq_b), _ptr(seg_b), _ptr(n_uniq), max_uniq, 1.0,
                                              _ptr(rows), _ptr(G), grid, nst, big) > 0
    G_ref = O.scatter_add_rows(dO, idx, N, 0) if R else np.zeros((N, D), np.float32)
    np.testing.assert_array_equal(G, G_ref)
    np.testing.assert_array_equal(rows[:U], G_ref[uniq])
    assert np.isnan(rows[U:]).all()                    # rows beyond n_uniq are never written
    rows2 = np.full((max_uniq, D), np.nan, np.float32)
    emu_rows.emu_scatter_add_rows_ring(_ptr(dOc), D, gr, _ptr(perm), _ptr(uniq_b), _ptr(seg_b), _ptr(n_uniq), max_uniq, 0.25,
                                       _ptr(rows2), None, grid, nst, big)
    np.testing.assert_array_equal(rows2[:U], G_ref[uniq] * np.float32(0.25))


@pytest.mark.parametrize("N,R,pad,digit", [(50, 300, 0, 0), (97001, 5000, 0, 0), (1000, 4100, 0, 5), (300, 2048, None, 0), (7, 40, 0, 2),
                                           (5000, 2049, 0, 0), (64, 100, 0, 0), (20, 0, 0, 0), (9, 30, 3, 0)])
def test_scatter_plan_emulated(emu_rows, N, R, pad, digit):
    """pr_scatter_plan's kernels (rows_plan.cuh: radix passes with the id conversion folded into the first histogram, run flags
    folded into the scan, emission folded into its apply pass) == a stable sort of (id, position): perm, the distinct ids, the run
    boundaries, n_uniq and row2slot -- with padding ids, out-of-range ids (flagged, dropped), several tiles and digit widths."""
    g = np.random.default_rng(N + R)
    idx = g.integers(0, N, size=R).astype(np.int64)
    if R > 10:
        idx[3] = -5                       # out of range: dropped + status bit 0
        idx[7] = N + 2
    padv = -1 if pad is None else pad
    key = np.where((idx < 0) | (idx >= N) | (idx == padv), N, idx)
    perm_ref = np.argsort(key, kind="stable").astype(np.int32)
    sk = key[perm_ref]
    nreal = int((sk < N).sum())
    starts = np.flatnonzero(np.r_[True, sk[1:nreal] != sk[:nreal - 1]]) if nreal else np.zeros(0, np.int64)
    U = len(starts)
    cap = max(1, min(R, N))
    perm = np.full(max(R, 1), -1, np.int32)
    uniq = np.full(cap, -1, np.int32)
    seg = np.full(cap + 1, -1, np.int32)
    n_uniq = np.full(1, -1, np.int32)
    row2slot = np.full(N, -1, np.int32)
    status = np.zeros(1, np.int32)
    emu_rows.emu_scatter_plan(_ptr(idx), R, N, padv, _ptr(perm), _ptr(uniq), _ptr(seg), _ptr(n_uniq), _ptr(row2slot), _ptr(status), digit)
    assert n_uniq[0] == U
    assert seg[U] == nreal and np.array_equal(seg[:U], starts)
    if R:
        assert np.array_equal(perm[:R], perm_ref)
        assert np.array_equal(uniq[:U], sk[starts])
        ref_slot = np.full(N, -1, np.int32)
        ref_slot[sk[starts]] = np.arange(U)
        assert np.array_equal(row2slot, ref_slot)
        assert status[0] == (1 if R > 10 else 0)
