"""GPU parity: table row kernels (K1 gather, K2 scatter-add, K10 AdamW) through the C ABI vs the oracle.
Integer / index work is bit-exact; the scatter-add is bit-exact too because the oracle defines the same
summation order."""
import numpy as np
import pytest
import torch

from oracle import sasrec_np as O
from tests.gpu_util import dev, rel, t, zipf_ids

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("impl", [1, 2, 0])
@pytest.mark.parametrize("N,D,shape", [(1001, 128, (7, 2, 11)), (5003, 512, (64, 2, 21)), (300, 64, (5,)), (97, 2048, (33,)),
                                      (50, 4, (9,)), (2000, 512, (4096, 2, 21))])
def test_gather_bit_exact(impl, N, D, shape):
    from pixelrec_b200 import ops
    g = np.random.default_rng(N + D)
    W = g.standard_normal((N, D)).astype(np.float32)
    idx = g.integers(0, N, size=shape).astype(np.int64)
    idx.reshape(-1)[:3] = [0, N - 1, 0]
    if impl == 2 and D * 4 < 16:
        pytest.skip("bulk path needs >= 16-byte rows")
    out = ops.gather_rows(t(W), t(idx), impl=impl)
    torch.cuda.synchronize()
    assert out.shape == shape + (D,)
    assert np.array_equal(out.cpu().numpy(), O.gather_rows(W, idx))


@pytest.mark.parametrize("impl", [1, 2])
def test_gather_empty_and_out_of_range(impl):
    from pixelrec_b200 import ops
    d = dev()
    W = torch.randn(10, 128, device=d)
    out = ops.gather_rows(W, torch.zeros(0, dtype=torch.int64, device=d), impl=impl)
    assert out.shape == (0, 128)
    status = torch.zeros(1, dtype=torch.int32, device=d)
    idx = torch.tensor([1, 10, 3, -1, 9], device=d)
    out = ops.gather_rows(W, idx, impl=impl, status=status)
    torch.cuda.synchronize()
    assert status.item() == 1                       # reference: IndexError (CPU) / device assert (CUDA)
    assert torch.equal(out[[0, 2, 4]], W[[1, 3, 9]]) and out[1].abs().sum() == 0 and out[3].abs().sum() == 0
    with pytest.raises(Exception):
        ops.gather_rows(W.cpu(), idx.cpu())          # no CPU fallback


def _plan_np(plan):
    U = int(plan.n_uniq.item())
    return U, plan.uniq_ids[:U].cpu().numpy(), plan.seg_start[:U + 1].cpu().numpy(), plan.perm[:plan.R].cpu().numpy()


@pytest.mark.parametrize("N,R,pad", [(50, 1, 0), (50, 37, 0), (1001, 2048, 0), (1001, 2049, 0), (97001, 5000, 0),
                                     (300, 70000, 0), (70000, 300000, 0), (1 << 17, 4097, 0), (40, 1000, None),
                                     (257, 999, 5)])
def test_scatter_plan_sorted_segments(N, R, pad):
    from pixelrec_b200 import ops
    g = np.random.default_rng(R)
    idx = g.integers(0, N, size=R).astype(np.int64)
    idx[g.random(R) < 0.3] = 0 if pad is None else pad       # lots of padding rows
    plan = ops.ScatterPlan(t(idx), N, pad)
    torch.cuda.synchronize()
    U, uniq, seg, perm = _plan_np(plan)
    keep = np.ones(R, bool) if pad is None else (idx != pad)
    eu, ec = np.unique(idx[keep], return_counts=True)
    assert U == len(eu) and np.array_equal(uniq, eu)
    assert seg[0] == 0 and np.array_equal(np.diff(seg), ec)
    # stable: inside a run positions ascend; runs hold exactly the positions of that id
    order = np.argsort(idx + (~keep) * (N + 1), kind="stable")
    assert np.array_equal(perm[:seg[-1]], order[:seg[-1]])
    assert sorted(perm.tolist()) == list(range(R))


def test_scatter_plan_all_padding_and_empty():
    from pixelrec_b200 import ops
    d = dev()
    plan = ops.ScatterPlan(torch.zeros(100, dtype=torch.int64, device=d), 10, 0)
    assert plan.n_uniq.item() == 0 and plan.seg_start[0].item() == 0
    plan = ops.ScatterPlan(torch.zeros(0, dtype=torch.int64, device=d), 10, 0)
    assert plan.n_uniq.item() == 0
    status = torch.zeros(1, dtype=torch.int32, device=d)
    plan = ops.ScatterPlan(torch.tensor([3, 11, 3, -2], device=d), 10, 0, status=status)
    assert status.item() == 1 and plan.n_uniq.item() == 1 and plan.uniq_ids[0].item() == 3


@pytest.mark.parametrize("N,D,R,zipf", [(101, 128, 500, False), (97001, 512, 42 * 256, True), (2001, 64, 3000, True),
                                        (501, 2048, 700, False), (64, 4096, 300, False), (1001, 256, 1, False),
                                        (33, 512, 5000, False)])
def test_scatter_add_bit_exact_vs_oracle(N, D, R, zipf):
    from pixelrec_b200 import ops
    g = np.random.default_rng(D + R)
    idx = zipf_ids(g, R, N) if zipf else g.integers(0, N, size=R).astype(np.int64)
    idx[g.random(R) < 0.25] = 0
    dO = g.standard_normal((R, D)).astype(np.float32)
    G_ref = O.scatter_add_rows(dO, idx, N, 0)
    plan = ops.ScatterPlan(t(idx), N, 0)
    G = torch.full((N, D), 7.0, device=dev())                     # NOT zero-filled: only touched rows are written
    rows = ops.scatter_add_rows(t(dO), plan, dense_G=G)
    torch.cuda.synchronize()
    U, uniq, seg, perm = _plan_np(plan)
    assert np.array_equal(rows[:U].cpu().numpy(), G_ref[uniq])    # bit-exact (same summation order)
    Gn = G.cpu().numpy()
    assert np.array_equal(Gn[uniq], G_ref[uniq])
    untouched = np.setdiff1d(np.arange(N), uniq)
    assert (Gn[untouched] == 7.0).all() and (G_ref[untouched] == 0).all()
    assert 0 not in uniq                                          # padding row never receives gradient
    # scale (DDP-mean / loss scaling) is applied after the sum
    rows2 = ops.scatter_add_rows(t(dO), plan, scale=0.125)
    assert np.array_equal(rows2[:U].cpu().numpy(), G_ref[uniq] * np.float32(0.125))


def test_scatter_linearity_full_size():
    """Pixel200K-shape (C2) sized property check: scatter(a) + scatter(b) == scatter(a+b) within fp32
    rounding, and column sums are preserved (sum of all gradient rows == sum of non-pad dOut rows)."""
    from pixelrec_b200 import ops
    N, D, B, L = 97001, 512, 1024, 20
    g = np.random.default_rng(0)
    idx = zipf_ids(g, B * 2 * (L + 1), N)
    idx[g.random(idx.size) < 0.2] = 0
    ti = t(idx)
    a = torch.randn(idx.size, D, device=dev())
    plan = ops.ScatterPlan(ti, N, 0)
    ra = ops.scatter_add_rows(a, plan)
    U = plan.n_uniq.item()
    keep = (ti != 0)
    assert torch.allclose(ra[:U].sum(0, dtype=torch.float64), a[keep].sum(0, dtype=torch.float64), rtol=1e-6, atol=1e-3)
    assert U == len(np.unique(idx[idx != 0]))


@pytest.mark.parametrize("N,D", [(257, 128), (1001, 512), (64, 2048), (100, 36)])
def test_adamw_rows_vs_oracle(N, D):
    from pixelrec_b200 import ops
    g = np.random.default_rng(N)
    W = (0.02 * g.standard_normal((N, D))).astype(np.float32)
    M = np.zeros_like(W)
    V = np.zeros_like(W)
    dW, dM, dV = t(W.copy()), t(M.copy()), t(V.copy())
    row2slot = torch.full((N,), -1, dtype=torch.int32, device=dev())
    for step in (1, 2, 3):
        R = 3 * N // 2
        idx = g.integers(0, N, size=R).astype(np.int64)
        idx[::3] = 0
        dO = (0.01 * g.standard_normal((R, D))).astype(np.float32)
        G = O.scatter_add_rows(dO, idx, N, 0)
        W, M, V = O.adamw_step(W, G, M, V, step, 1e-3, 0.1)
        plan = ops.ScatterPlan(t(idx), N, 0, row2slot=row2slot)
        rows = ops.scatter_add_rows(t(dO), plan)
        ops.adamw_rows(dW, dM, dV, rows, row2slot, 1e-3, 0.9, 0.999, 1e-8, 0.1, step)
        torch.cuda.synchronize()
        assert (row2slot == -1).all()                              # consumed entries are restored
        assert np.abs(dW.cpu().numpy() - W).max() < 3e-7
        assert rel(dM.cpu().numpy(), M) < 1e-5 and rel(dV.cpu().numpy(), V, 1e-12) < 1e-4
    # rows without gradient still decay (dense AdamW semantics of trainer.py:102)
    assert np.abs(dW.cpu().numpy()[0] - W[0]).max() < 3e-7 and not np.array_equal(W[0], (0.02 * np.ones(1)))


def test_adamw_rows_matches_golden(golden):
    """Two AdamW steps on the table with the reference's own gradient (goldens from torch.optim.AdamW)."""
    from pixelrec_b200 import ops
    key = "item_embedding.weight"
    W = t(golden["params"][key].copy())
    G = golden["grads"][key]
    N, D = G.shape
    M, V = torch.zeros_like(W), torch.zeros_like(W)
    nz = np.flatnonzero(np.abs(G).sum(1) > 0)
    row2slot = torch.full((N,), -1, dtype=torch.int32, device=dev())
    rows = t(G[nz])
    for step, gkey in ((1, "adamw1/" + key), (2, "adamw2/" + key)):
        row2slot[t(nz)] = torch.arange(len(nz), dtype=torch.int32, device=dev())
        ops.adamw_rows(W, M, V, rows, row2slot, 1e-4, 0.9, 0.999, 1e-8, 0.1, step)
        assert np.abs(W.cpu().numpy() - golden[gkey]).max() < 3e-7


@pytest.mark.parametrize("n", [1, 3, 4, 1023, 100003])
def test_adamw_dense_vs_oracle(n):
    from pixelrec_b200 import ops
    g = np.random.default_rng(n)
    w = g.standard_normal(n).astype(np.float32)
    gr = g.standard_normal(n).astype(np.float32)
    m = np.zeros_like(w)
    v = np.zeros_like(w)
    dw, dg, dm, dv = t(w.copy()), t(gr), t(m.copy()), t(v.copy())
    step_dev = torch.zeros(1, dtype=torch.int64, device=dev())
    for step in (1, 2):
        w, m, v = O.adamw_step(w, gr, m, v, step, 1e-3, 0.01)
        step_dev.fill_(step)
        ops.adamw_dense(dw, dg, dm, dv, 1e-3, 0.9, 0.999, 1e-8, 0.01, 0, step_dev=step_dev)   # step read from device
        assert np.abs(dw.cpu().numpy() - w).max() < 1e-6


@pytest.mark.parametrize("n_seq,L,N,B", [(50, 10, 40, 64), (500, 20, 5000, 300), (7, 5, 8, 20)])
def test_on_device_batch_builder_bit_exact(n_seq, L, N, B):
    """A1: pr_seq_batch_build == oracle restatement (same Philox stream), and the batch has SEQTrainDataset's layout:
    negatives never collide with the sequence, neg/mask are 0 up to and including the first real item."""
    from pixelrec_b200 import ops
    g = np.random.default_rng(n_seq + B)
    W = L + 1
    padded = np.zeros((n_seq, W), dtype=np.int64)
    for i in range(n_seq):
        n = int(g.integers(1, W + 1))
        padded[i, W - n:] = g.integers(1, N, size=n)
    sel = g.integers(0, n_seq, size=B).astype(np.int64)
    ref_items, ref_mask = O.seq_batch_build(padded, sel, N, 777)
    items, mask = ops.seq_batch_build(t(padded), t(sel), N, 777)
    assert np.array_equal(items.cpu().numpy(), ref_items) and np.array_equal(mask.cpu().numpy(), ref_mask)
    it, mk = ref_items, ref_mask
    for b in range(B):
        pos, neg = it[b, 0], it[b, 1]
        n = int((pos != 0).sum())
        assert mk[b].sum() == max(n - 1, 0) and (neg != 0).sum() == max(n - 1, 0)
        assert not (set(neg[neg != 0].tolist()) & set(pos[pos != 0].tolist()))
        assert ((neg >= 0) & (neg < N)).all()
    other, _ = ops.seq_batch_build(t(padded), t(sel), N, 778)
    assert not torch.equal(other, items)                      # a new seed draws new negatives
