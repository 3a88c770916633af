"""CUDA-graph replay of the training step (staged: written without GPU access, opt-in through yaml `cuda_graph: True`).

At the reference's shipped batch size (train_batch_size: 64, overall/ID.yaml) a step is ~125 kernel launches of a few
microseconds each and the host cannot enqueue them fast enough (2.3 ms/step measured).  The step
(REC/trainer/trainer.py:116-125: zero_grad -> forward -> backward -> optimizer.step) contains no host synchronisation here, so
it can be captured once and replayed with one launch.  Two host-side scalars would otherwise be frozen into the graph:

  * the dropout seed of the forward call -> pr_set_seed_device: the kernels add a device-resident offset that a captured
    `add_(1)` bumps at the end of every replay, so replay j uses exactly the seed eager step j would have used;
  * AdamW's step count (bias correction) -> FusedAdamW.use_device_step(): read from device memory, bumped inside the graph.

The caller must not hold the loss tensor (or anything else with a grad_fn) of an earlier EAGER step when the capture starts:
a live autograd graph keeps its AccumulateGrad nodes, which stay bound to the stream they were created on (the default one),
and the engine would then make that stream wait on the capturing one -- cudaErrorStreamCaptureImplicit.

Multi-GPU: capturable with the peer-memory exchange (`exchange: p2p`, the default) whose index plan and barrier epochs stay on
the device (dist.PeerPlanDev, pr_peer_barrier); the NCCL all_to_all exchange needs host-side split sizes and stays eager.
Fixed batch shape (a ragged last batch runs eagerly); dropout > 0 needs device-side seeds (-DPR_SEED_DEV, the default build).
"""
import torch

from .. import ops


class GraphedTrainStep:
    def __init__(self, model, optimizer, example_batch):
        dev = example_batch[0].device
        self.model, self.optimizer = model, optimizer
        self.static = tuple(torch.empty_like(t) for t in example_batch)
        for s, t in zip(self.static, example_batch):
            s.copy_(t)
        self.seed_offset = torch.zeros(1, dtype=torch.int64, device=dev)
        needs_seed = model.training and any(getattr(model, k, 0.0) for k in ("hidden_dropout_prob", "attn_dropout_prob",
                                                                              "dropout_prob"))
        self._seeded = bool(needs_seed)
        if self._seeded:
            ops.set_seed_device(self.seed_offset)            # raises in builds without -DPR_SEED_DEV
        optimizer.use_device_step(True)
        # row-sharded table (N > 1): only the peer-memory exchange with the device-side plan is free of host syncs; its barrier
        # epochs move to device memory for the replays
        self._px = []
        for m in model.modules():
            if hasattr(m, "graph_capturable"):
                if not m.graph_capturable():
                    optimizer.use_device_step(False)
                    raise ops._lib.PixelRecB200Error("cuda_graph: this table's row exchange synchronises with the host "
                                                     "(needs exchange='p2p' with the device plan, after one eager step)")
                m._px.use_device_epoch(True)
                self._px.append(m._px)
        self.graph = torch.cuda.CUDAGraph()
        torch.cuda.synchronize(dev)
        launches0 = ops.LAUNCHES["count"]
        try:
            with torch.cuda.graph(self.graph):
                optimizer.zero_grad()
                self.loss = model(self.static)
                self.loss.backward()
                optimizer.step()
                if self._seeded:
                    self.seed_offset.add_(1)
        except Exception as e:
            self.close()
            if "legacy stream depend on a capturing" in str(e):
                raise RuntimeError("CUDA-graph capture of the training step failed because an autograd graph of an earlier eager "
                                   "step is still alive (drop the previous loss tensor before capturing)") from e
            raise
        self.launches_per_replay = ops.LAUNCHES["count"] - launches0     # kernels of ours inside the graph
        ops.LAUNCHES["count"] = launches0                                 # capture itself launched nothing
        # capturing ran the step's Python once (seed counter, optimizer._step) without executing a kernel: the first replay IS that step
        self._first = True

    def matches(self, batch):
        return len(batch) == len(self.static) and all(b.shape == s.shape and b.dtype == s.dtype for b, s in zip(batch, self.static))

    def __call__(self, batch):
        for s, b in zip(self.static, batch):
            s.copy_(b, non_blocking=True)
        if not self._first:
            self.optimizer._step += 1            # host mirror of the device-side count (checkpoints)
            rng = getattr(self.model, "rng", None)
            if rng is not None:
                rng.calls += 1                   # host mirror of the device-side seed offset
        self._first = False
        ops.LAUNCHES["count"] += self.launches_per_replay
        self.graph.replay()
        return self.loss

    def close(self):
        """back to eager execution: host-side seeds and step count take over where the replays stopped"""
        if self._seeded:
            ops.set_seed_device(None)
        self.optimizer.use_device_step(False)
        for px in getattr(self, "_px", []):
            px.use_device_epoch(False)
        self._px = []
        # release the captured graph now: it holds references to NCCL work (N > 1) and to the private memory pool, and a process
        # group cannot be destroyed cleanly while such a graph is alive
        self.graph = None
        self.loss = None
