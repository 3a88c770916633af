"""N>1 path on CPU: world_size-2 gloo processes exercise the row-sharded table's host logic (index bucketing,
all_to_all exchange, un-permutation, owner-side reduction, state-dict gather/scatter) with the oracle standing in
for the CUDA row kernels.  Checked against the single-process dense oracle."""
import os
import socket
import types

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import sasrec_np as O


class OracleRows:
    """CPU stand-in for pixelrec_b200.dist.CudaRows built on the oracle (test infrastructure only)."""

    @staticmethod
    def gather(W, idx):
        return torch.from_numpy(O.gather_rows(W.detach().numpy(), idx.numpy()))

    @staticmethod
    def plan(idx, N, padding_idx, row2slot=None):
        return types.SimpleNamespace(idx=idx.numpy().copy(), N=N, pad=-1 if padding_idx is None else padding_idx)

    @staticmethod
    def scatter(dOut, plan):
        G = O.scatter_add_rows(dOut.numpy(), plan.idx, plan.N, plan.pad)
        return torch.from_numpy(G)                 # dense over the local shard (the CUDA path returns sparse rows)

    @staticmethod
    def scatter_slots(dOut, slot_idx, U, pad_slot):
        return torch.from_numpy(O.scatter_add_rows(dOut.numpy(), slot_idx.numpy(), U, pad_slot))


class ShmBuf:
    """CPU stand-in for ops.SharedBuffer: a /dev/shm file both gloo processes map (test infrastructure only)."""

    def __init__(self, nbytes, device):
        self.path = f"/dev/shm/pr_b200_test_{os.getpid()}_{id(self)}"
        self.mm = np.memmap(self.path, dtype=np.uint8, mode="w+", shape=(int(nbytes),))
        self.ref = torch.from_numpy(self.mm)
        self.handle = self.path.encode()

    def tensor(self, shape, dtype):
        n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        return self.ref[:n].view(dtype).view(*shape)


class ShmPeer:
    """CPU stand-in for pixelrec_b200.dist.CudaPeer: same slot protocol as csrc/peer.cu, python loops."""
    alloc = ShmBuf

    @staticmethod
    def open(handle, device):
        return torch.from_numpy(np.memmap(handle.decode(), dtype=np.uint8, mode="r+"))

    @staticmethod
    def table(refs, device):
        return list(refs)

    @staticmethod
    def gather(shard_table, G, N, D, idx):
        out = torch.empty(idx.numel(), D)
        for r, i in enumerate(idx.reshape(-1).tolist()):
            assert 0 <= i < N
            out[r] = shard_table[i % G].view(torch.float32).view(-1, D)[i // G]
        return out.view(*idx.shape, D)

    @staticmethod
    def barrier(flag_table, G, rank, epoch, status):
        """same flag protocol as peer_barrier_kernel: publish `epoch` in every peer's array, wait for every peer's flag"""
        import time
        for r in range(G):
            flag_table[r].view(torch.int64)[rank] = epoch
        t0 = time.time()
        mine = flag_table[rank].view(torch.int64)
        while any(int(mine[r]) < epoch for r in range(G)):
            if time.time() - t0 > 60:
                status[0] |= 4
                return
            time.sleep(0.001)

    @staticmethod
    def push(rows, ids, G, rank, cap, skip_id, rows_table, ids_table, counters, status):
        for u, i in enumerate(ids.tolist()):
            if i == skip_id:
                continue
            o = i % G
            pos = int(counters[o])
            counters[o] += 1
            if pos >= cap:
                status[0] |= 2
                continue
            slot = rank * cap + pos
            rows_table[o].view(torch.float32).view(-1, rows.shape[1])[slot] = rows[u]
            ids_table[o].view(torch.int64)[slot] = i // G


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, N, D, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pixelrec_b200 import dist as pd
        pd.ROWS = OracleRows
        g = np.random.default_rng(123)
        W = g.standard_normal((N, D)).astype(np.float32)
        idx_all = g.integers(0, N, size=(world, 6, 2, 5)).astype(np.int64)
        idx_all[:, :, 1, 0] = 0
        idx_all[0, 0, 0, :] = [0, 1, 2, 3, N - 1]
        dE_all = g.standard_normal((world, 6, 2, 5, D)).astype(np.float32)
        table = pd.ShardedTableEmbedding(N, D, padding_idx=0, exchange="nccl")
        table.sink.row2slot = torch.zeros(1)            # oracle backend ignores it
        assert table.n_local == len(range(rank, N, world))
        table.load_state_dict({"weight": torch.from_numpy(W)})
        assert torch.equal(table.weight.detach(), torch.from_numpy(W[rank::world]))
        idx = torch.from_numpy(idx_all[rank])
        E = table(idx)
        assert torch.equal(E.detach(), torch.from_numpy(W[idx_all[rank]]))           # bit-exact gather across shards
        E.backward(torch.from_numpy(dE_all[rank]))
        (splan, G_local), = table.sink.pending
        # single-process dense oracle over the union of both ranks' lookups
        G_full = O.scatter_add_rows(dE_all.reshape(-1, D), idx_all.reshape(-1), N, 0)
        ref = G_full[rank::world]
        assert np.allclose(G_local.numpy(), ref, rtol=1e-5, atol=1e-6)
        if rank == 0:
            assert (G_local.numpy()[0] == 0).all()                                     # global pad id 0 lives here
        full = table.full_weight()
        assert torch.equal(full, torch.from_numpy(W))
        sd = table.state_dict()
        assert sd["weight"].shape == (N, D) and torch.equal(sd["weight"], torch.from_numpy(W))
        out_q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        out_q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def _worker_p2p(rank, world, port, N, D, cap_factor, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["PR_P2P_CAP_FACTOR"] = str(cap_factor)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pixelrec_b200 import dist as pd
        pd.ROWS = OracleRows
        pd.PEER = ShmPeer
        g = np.random.default_rng(321)
        W = g.standard_normal((N, D)).astype(np.float32)
        table = pd.ShardedTableEmbedding(N, D, padding_idx=0, exchange="p2p")
        table.sink.row2slot = torch.zeros(1)            # oracle backend ignores it
        table.load_state_dict({"weight": torch.from_numpy(W)})
        Wcur = W.copy()
        for step in range(3):                           # several steps: slot reset, barriers, peers see updated rows
            idx_all = g.integers(0, N, size=(world, 6, 2, 5)).astype(np.int64)
            idx_all[:, :, 1, 0] = 0
            idx_all[0, 0, 0, :] = [0, 1, 2, 3, N - 1]
            idx_all[1, 1, 0, :2] = idx_all[0, 2, 0, :2]                              # same ids requested from both ranks
            dE_all = g.standard_normal((world, 6, 2, 5, D)).astype(np.float32)
            idx = torch.from_numpy(idx_all[rank])
            if step == 1:                               # plan built ahead of time, as trainer.Lookahead does
                plan = table._make_plan(idx)
                assert isinstance(plan, pd.PeerPlan) and plan.U == len(np.unique(idx_all[rank]))
            E = table(idx)
            assert torch.equal(E.detach(), torch.from_numpy(Wcur[idx_all[rank]]))    # bit-exact gather across shards
            E.backward(torch.from_numpy(dE_all[rank]))
            (splan, G_local), = table.sink.pending
            table.sink.pending.clear()
            G_full = O.scatter_add_rows(dE_all.reshape(-1, D), idx_all.reshape(-1), N, 0)
            assert np.allclose(G_local.numpy(), G_full[rank::world], rtol=1e-5, atol=1e-6)
            if rank == 0:
                assert (G_local.numpy()[0] == 0).all()                               # global pad id 0 lives here
            assert table.exchange_status() == 0
            assert (table._px.recv_ids == -1).all()                                  # receive slots released
            with torch.no_grad():                       # stand-in for the optimizer step on the owned shard
                table.weight.data -= 0.5 * G_local
            Wnew = table.full_weight().numpy()          # all_gather path: independent of the peer-memory reads
            assert np.allclose(Wnew, Wcur - 0.5 * G_full, rtol=1e-5, atol=1e-6)
            Wcur = Wnew.copy()
        assert torch.equal(table.full_weight(), torch.from_numpy(Wcur))
        sd = table.state_dict()
        assert torch.equal(sd["weight"], torch.from_numpy(Wcur))
        dist.barrier()
        out_q.put((rank, "ok"))
    except Exception:  # pragma: no cover
        import traceback
        out_q.put((rank, traceback.format_exc()))
    finally:
        px = getattr(locals().get("table"), "_px", None)
        if px is not None:
            for b in (px._w, px._rows, px._ids):
                try:
                    os.unlink(b.path)
                except OSError:
                    pass
        dist.destroy_process_group()


def _worker_p2p_overflow(rank, world, port, out_q):
    """a receive region too small for the step: rows are dropped, never written out of bounds, and the flag is raised"""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), PR_P2P_CAP_FACTOR="0.01")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pixelrec_b200 import dist as pd
        pd.ROWS = OracleRows
        pd.PEER = ShmPeer
        N, D = 4001, 4
        table = pd.ShardedTableEmbedding(N, D, padding_idx=0, exchange="p2p")
        table.sink.row2slot = torch.zeros(1)
        idx = torch.arange(1 + rank, 1 + rank + 2 * 400, 2).view(20, 20)    # 400 distinct ids, all owned by ONE rank; cap = 64
        E = table(idx)
        E.backward(torch.ones_like(E))
        assert table._px.cap == 64
        st = torch.tensor([table.exchange_status()])
        dist.all_reduce(st, op=dist.ReduceOp.MAX)
        assert int(st) == 2
        dist.barrier()
        out_q.put((rank, "ok"))
    except Exception:  # pragma: no cover
        import traceback
        out_q.put((rank, traceback.format_exc()))
    finally:
        px = getattr(locals().get("table"), "_px", None)
        if px is not None:
            for b in (px._w, px._rows, px._ids):
                try:
                    os.unlink(b.path)
                except OSError:
                    pass
        dist.destroy_process_group()


def _spawn(target, world, *args):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=target, args=(r, world, port) + args + (q,)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(30)
    assert all(r[1] == "ok" for r in res), res


@pytest.mark.parametrize("N,D,cap_factor", [(11, 8, 2.0), (64, 16, 2.0), (257, 4, 1.2)])
def test_sharded_table_peer_exchange_world2_gloo(N, D, cap_factor):
    """exchange='p2p' host logic (plans, slot protocol, barriers, buffer reuse over several steps) with /dev/shm standing
    in for peer-mapped HBM and the oracle for the row kernels"""
    _spawn(_worker_p2p, 2, N, D, cap_factor)


def test_peer_exchange_overflow_is_flagged_world2_gloo():
    _spawn(_worker_p2p_overflow, 2)


@pytest.mark.parametrize("N,D", [(11, 8), (64, 16)])
def test_sharded_table_world2_gloo(N, D):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, N, D, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(30)
    assert all(r[1] == "ok" for r in res), res


def test_shard_rows_partition():
    from pixelrec_b200.dist import shard_rows
    for N in (1, 7, 97001, 408375):
        for world in (1, 2, 4, 8):
            assert sum(shard_rows(N, world, r) for r in range(world)) == N


def _init_worker(rank, world, port, N, D, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pixelrec_b200 import dist as pd
        torch.manual_seed(2020)                      # run.py seeds every rank identically (init_seed)
        table = pd.ShardedTableEmbedding(N, D, padding_idx=0)
        table.init_normal_(0.0, 0.02, chunk_rows=7)  # small chunks: the block / shard arithmetic is exercised
        full = table.full_weight()                   # collective
        moment = torch.arange(table.n_local * D, dtype=torch.float32).view(table.n_local, D) + 1000 * rank
        m_full = table.gather_rows_full(moment)      # Adam-moment layout used by FusedAdamW.state_dict()
        back = table.shard_of_full(m_full)
        after = torch.rand(1)                        # generators of all ranks stay in lockstep
        if rank == 0:
            out_q.put(("ok", full.numpy(), torch.equal(back, moment), float(after)))
        else:
            out_q.put(("ok1", None, torch.equal(back, moment), float(after)))
    except Exception:  # pragma: no cover
        import traceback
        out_q.put(("err", traceback.format_exc(), None, None))
    finally:
        dist.destroy_process_group()


def test_sharded_table_init_is_the_logical_table_world2_gloo():
    """with one seed on every rank, `weight.normal_()` per shard gave up to `world` items the same embedding.  The shard
    must be rows rank::world of the logical [N, D] table: all rows distinct, iid, independent of the world size; and the
    optimizer-state gather / re-shard round trip is exact."""
    N, D, world = 45, 8, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_init_worker, args=(r, world, port, N, D, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(60)
    assert all(r[0].startswith("ok") for r in res), res
    full = next(r[1] for r in res if r[0] == "ok")
    assert all(r[2] for r in res)                                   # gather_rows_full -> shard_of_full is the identity
    assert res[0][3] == res[1][3]                                   # same random stream position on both ranks
    assert full.shape == (N, D) and len(np.unique(full.round(7), axis=0)) == N      # no two items share an embedding
    # the same seed in ONE process (world 1) draws the same logical table
    from pixelrec_b200 import dist as pd
    torch.manual_seed(2020)
    single = pd.ShardedTableEmbedding(N, D, padding_idx=0)
    single.init_normal_(0.0, 0.02, chunk_rows=7)
    assert np.array_equal(single.weight.detach().numpy(), full)


def _cpu_score_fn(seq_all, W_local, k, hu, hi, mask_col0):
    """stand-in for ops.score_topk_exact on the CPU: fp32 scores, pad column / history masked, (score desc, id asc) order"""
    from pixelrec_b200.dist import merge_topk_candidates
    s = seq_all @ W_local.t()
    if mask_col0:
        s[:, 0] = -float("inf")
    if hu is not None:
        s[hu, hi] = -float("inf")
    ids = torch.arange(W_local.shape[0]).expand_as(s).contiguous()
    return merge_topk_candidates(s, ids, k)


def _worker_sharded_topk(rank, world, port, N, D, B_e, k, out_q):
    try:
        import torch.distributed as dist
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
        from pixelrec_b200.dist import ShardedTopK, merge_topk_candidates
        g = torch.Generator().manual_seed(5)                      # same on every rank: the logical problem
        W = torch.randint(-3, 4, (N, D), generator=g).float()     # integers: fp32 scores are exact, ties are frequent
        seq = torch.randint(-3, 4, (world, B_e, D), generator=g).float()
        n_hist = [int(x) for x in torch.randint(0, 3 * B_e, (world,), generator=g)]
        n_hist[-1] = 0                                            # one rank without history pairs
        hist = [(torch.randint(0, B_e, (n,), generator=g), torch.randint(1, N, (n,), generator=g)) for n in n_hist]
        scorer = ShardedTopK(world, rank, score_fn=_cpu_score_fn)
        for _ in range(2):                                        # reusable across batches
            val, idx = scorer(seq[rank], W[rank::world].contiguous(), k, hist[rank][0], hist[rank][1], pad_id=0)
            full = seq[rank] @ W.t()
            full[:, 0] = -float("inf")
            full[hist[rank][0], hist[rank][1]] = -float("inf")
            ref_v, ref_i = merge_topk_candidates(full, torch.arange(N).expand_as(full).contiguous(), k)
            assert torch.equal(idx, ref_i), (idx[:2], ref_i[:2])
            assert torch.equal(val, ref_v)
        dist.destroy_process_group()
        out_q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        import traceback
        out_q.put((rank, "fail: " + repr(e) + traceback.format_exc()))


@pytest.mark.parametrize("world,N,D,B_e,k", [(2, 101, 8, 7, 10), (3, 64, 4, 5, 3)])
def test_sharded_topk_merge_equals_full_catalog_ranking_gloo(world, N, D, B_e, k):
    """ShardedTopK: all-gather of the encoder outputs and history pairs, per-shard candidates, all-to-all + merge == the ranking of
    the whole catalog (ids AND order, ties by lower id), with uneven shards and a rank that has no history pairs."""
    _spawn(_worker_sharded_topk, world, N, D, B_e, k)
