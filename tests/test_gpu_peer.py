"""GPU parity of the peer-memory row kernels (csrc/peer.cu) on ONE device: the "peers" are separate allocations of the same
GPU, so pr_gather_rows_peers_f32 / pr_push_rows_peers_f32 and the SharedBuffer aliasing are exercised without NVLink.
The multi-process path (CUDA IPC + NCCL barriers) is tests/test_gpu_dist.py[p2p]; its host logic runs on CPU in
tests/test_dist_gloo.py."""
import os

import numpy as np
import pytest
import torch

from oracle import sasrec_np as O
from tests.gpu_util import dev, t

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("G,N,D,R", [(1, 50, 8, 33), (2, 101, 64, 500), (4, 1003, 512, 4097), (8, 97001, 512, 20000), (3, 77, 36, 10)])
def test_gather_rows_peers_bit_exact(G, N, D, R):
    from pixelrec_b200 import ops
    d = dev()
    g = np.random.default_rng(G * 1000 + N)
    W = g.standard_normal((N, D)).astype(np.float32)
    bufs = [ops.SharedBuffer(max(len(range(r, N, G)), 1) * D * 4, d) for r in range(G)]
    for r, b in enumerate(bufs):
        n_r = len(range(r, N, G))
        if n_r:
            b.tensor((n_r, D), torch.float32).copy_(t(W[r::G]))
    table = torch.tensor([b.ref for b in bufs], dtype=torch.int64, device=d)
    idx = g.integers(0, N, size=R).astype(np.int64)
    idx[:3] = [0, N - 1, 0]
    out = ops.gather_rows_peers(table, G, N, D, t(idx))
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), O.gather_rows(W, idx))
    status = torch.zeros(1, dtype=torch.int32, device=d)
    bad = idx.copy()
    bad[1] = N
    out = ops.gather_rows_peers(table, G, N, D, t(bad), status=status)
    assert int(status.item()) == 1 and (out[1] == 0).all()


@pytest.mark.parametrize("G,N,D,U", [(2, 101, 64, 60), (4, 1003, 512, 700), (8, 97001, 512, 30000)])
def test_push_rows_peers_then_owner_reduce(G, N, D, U):
    """every 'rank' pushes its distinct ids; each owner's plan + scatter-add over its receive buffer equals the dense oracle"""
    from pixelrec_b200 import ops
    d = dev()
    g = np.random.default_rng(G + N)
    cap = U                                                    # cannot overflow
    rows_b = [ops.SharedBuffer(G * cap * D * 4, d) for _ in range(G)]
    ids_b = [ops.SharedBuffer(G * cap * 8, d) for _ in range(G)]
    recv_rows = [b.tensor((G * cap, D), torch.float32) for b in rows_b]
    recv_ids = [b.tensor((G * cap,), torch.int64) for b in ids_b]
    for x in recv_ids:
        x.fill_(-1)
    rows_table = torch.tensor([b.ref for b in rows_b], dtype=torch.int64, device=d)
    ids_table = torch.tensor([b.ref for b in ids_b], dtype=torch.int64, device=d)
    status = torch.zeros(1, dtype=torch.int32, device=d)
    dense = np.zeros((N, D), np.float64)
    for rank in range(G):
        ids = np.sort(g.choice(N, size=U, replace=False)).astype(np.int64)
        ids[0] = 0                                              # the padding id is skipped by the push
        rows = g.standard_normal((U, D)).astype(np.float32)
        dense[ids[1:]] += rows[1:]
        counters = torch.zeros(G, dtype=torch.int32, device=d)
        ops.push_rows_peers(t(rows), t(ids), G, rank, cap, 0, rows_table, ids_table, counters, status)
        torch.cuda.synchronize()
        cnt = counters.cpu().numpy()
        assert cnt.sum() == U - 1 and all(cnt[o] == ((ids[1:] % G) == o).sum() for o in range(G))
    assert int(status.item()) == 0
    for owner in range(G):
        n_local = len(range(owner, N, G))
        plan = ops.ScatterPlan(recv_ids[owner], n_local, None)
        red = ops.scatter_add_rows(recv_rows[owner], plan)
        torch.cuda.synchronize()
        nu = int(plan.n_uniq.item())
        got = np.zeros((n_local, D), np.float64)
        got[plan.uniq_ids[:nu].cpu().numpy()] = red[:nu].cpu().numpy()
        assert np.allclose(got, dense[owner::G], rtol=1e-5, atol=1e-6)


def test_push_rows_peers_overflow_flag():
    from pixelrec_b200 import ops
    d = dev()
    G, D, cap, U = 2, 8, 4, 40
    rows_b = [ops.SharedBuffer(G * cap * D * 4, d) for _ in range(G)]
    ids_b = [ops.SharedBuffer(G * cap * 8, d) for _ in range(G)]
    guard = [b.tensor((G * cap,), torch.int64) for b in ids_b]
    for x in guard:
        x.fill_(-1)
    rows_table = torch.tensor([b.ref for b in rows_b], dtype=torch.int64, device=d)
    ids_table = torch.tensor([b.ref for b in ids_b], dtype=torch.int64, device=d)
    status = torch.zeros(1, dtype=torch.int32, device=d)
    counters = torch.zeros(G, dtype=torch.int32, device=d)
    ids = np.arange(2, 2 + 2 * U, 2).astype(np.int64)          # all owned by rank 0
    ops.push_rows_peers(t(np.ones((U, D), np.float32)), t(ids), G, 1, cap, 0, rows_table, ids_table, counters, status)
    torch.cuda.synchronize()
    assert int(status.item()) == 2
    assert (guard[0][:cap] == -1).all() and (guard[0][cap:] >= 0).all() and (guard[1] == -1).all()   # only rank 1's region, in bounds
