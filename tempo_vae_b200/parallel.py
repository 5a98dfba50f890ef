"""Data-parallel training across the GPUs of one box: one process per GPU, gradients all-reduced in buckets over
NCCL (NVLink 5 / NVSwitch) while backward is still running.

The reference has no distributed code at all; this is the sharding SURVEY.md §8(e) derives from
the path itself: samples are independent in forward/backward (GroupNorm and attention are per-sample, the loss is
a per-sample sum divided by the local batch), so the only exchange step is the sum of the 26.2 M gradient elements.

How it overlaps. FusedAdamW keeps all gradients in ONE flat fp32 buffer in parameter order. The buffer is cut into
~25 MB buckets at parameter boundaries. The engine's backward writes each parameter's gradient exactly once and
calls `ENGINE.grad_ready_hook(param)` right after enqueueing the kernel that produced it; when the last expected
parameter of a bucket has reported, the bucket's all-reduce is issued (async) — NCCL's stream waits on the compute
stream up to that point only, so communication of late layers (decoder.conv_out comes first) overlaps the
remaining backward. `finish()` issues whatever is left and makes the compute stream wait for all buckets.
Gradients are summed; the 1/world factor is folded into the fused AdamW kernel (`grad_scale`), so the global
grad-norm clip sees the averaged gradient, identical on every rank.

The noise of the reparameterisation is keyed by the GLOBAL sample index (Philox counter), so a run is invariant
to the number of ranks.
"""
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def bind_to_gpu_numa(device_index: int) -> bool:
    """Pin the calling process to the CPU cores that are local to GPU `device_index` (NVML's ideal CPU affinity), so
    that pinned host buffers allocated afterwards land on the GPU's NUMA node and H2D copies of 8 ranks do not all
    cross the inter-socket link. Best effort: returns False when NVML or the affinity call is unavailable."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        ncpu = os.cpu_count() or 1
        words = (ncpu + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * i + b for i, w in enumerate(mask) for b in range(64) if (w >> b) & 1]
        cpus = [c for c in cpus if c < ncpu]
        if not cpus:
            return False
        os.sched_setaffinity(0, cpus)
        return True
    except Exception:  # noqa: BLE001
        return False


class GradBucketer:
    """Device-agnostic bucket bookkeeping over a flat gradient buffer (also exercised on CPU with gloo)."""

    def __init__(self, flat_grad: torch.Tensor, ranges: Sequence[Tuple[object, int, int]], bucket_bytes: int = 25 << 20,
                 group=None):
        self.flat = flat_grad
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        per = max(1, bucket_bytes // flat_grad.element_size())
        self.buckets: List[Tuple[int, int]] = []
        self.bucket_of = {}
        start, cur_end, members = None, 0, []
        self.members: List[List[int]] = []
        for key, s, e in ranges:
            if start is None:
                start = s
            members.append(id(key))
            cur_end = e
            if cur_end - start >= per:
                self._close(start, cur_end, members)
                start, members = None, []
        if members:
            self._close(start, cur_end, members)
        # the last bucket runs to the end of the (padded) flat buffer
        if self.buckets:
            s, _ = self.buckets[-1]
            self.buckets[-1] = (s, flat_grad.numel())
        self.expected: Optional[List[set]] = None   # learned from the first backward
        self._reset()

    def _close(self, s, e, members):
        b = len(self.buckets)
        self.buckets.append((s, e))
        self.members.append(list(members))
        for m in members:
            self.bucket_of[m] = b

    def _reset(self):
        self.seen = [set() for _ in self.buckets]
        self.launched = [False] * len(self.buckets)
        self.works = []

    # called by the engine when a parameter's gradient has been enqueued
    def ready(self, param):
        b = self.bucket_of.get(id(param))
        if b is None:
            return
        self.seen[b].add(id(param))
        if self.expected is not None and not self.launched[b] and self.seen[b] >= self.expected[b]:
            self._launch(b)

    def _launch(self, b):
        s, e = self.buckets[b]
        self.launched[b] = True
        if self.world > 1:
            if self.flat.is_cuda:   # weight gradients are produced on a second stream (model.ENGINE.wgrad_overlap)
                from .model import ENGINE
                ENGINE.sync_streams_for_collective(self.flat.device)
            self.works.append(dist.all_reduce(self.flat[s:e], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self):
        """Issue the remaining buckets, wait (stream-ordered on CUDA) for all of them, learn the expected sets."""
        for b in range(len(self.buckets)):
            if not self.launched[b]:
                self._launch(b)
        for w in self.works:
            w.wait()
        if self.expected is None:
            self.expected = [set(s) for s in self.seen]
        self._reset()


class DataParallel:
    """Wraps (model, FusedAdamW) for one-process-per-GPU data parallelism.

        dp = DataParallel(model, model.optimizer)          # after dist.init_process_group("nccl")
        loss, metrics = dp.get_loss(x_local)               # same return as model.get_loss
        dp.backward(loss); dp.step(max_grad_norm=1.0)
    """

    def __init__(self, model, optimizer, bucket_mb: float = 25.0, group=None):
        from .model import ENGINE
        from .optim import FusedAdamW
        if not isinstance(optimizer, FusedAdamW):
            raise TypeError("DataParallel needs the flat-buffer FusedAdamW optimiser")
        if not dist.is_initialized():
            raise RuntimeError("call torch.distributed.init_process_group first")
        self.model, self.optimizer, self.group = model, optimizer, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.engine = ENGINE
        # identical replicas: rank 0's parameters win (they are one flat buffer => one broadcast)
        dist.broadcast(optimizer.flat_param, src=0, group=group)
        ENGINE.params_changed()
        self.bucketer = GradBucketer(optimizer.flat_grad, optimizer.param_ranges(), int(bucket_mb * (1 << 20)), group)

    def get_loss(self, x, **kw):
        B = x.shape[0]
        return self.model.get_loss(x, sample_offset=self.rank * B, global_batch=self.world * B, **kw)

    def backward(self, loss, sync: bool = True):
        """sync=False: a micro-batch that is not the last of its optimiser step -- gradients are only accumulated
        into the flat buffer (no hooks, no collective); the all-reduce of the step runs, overlapped as usual, with the
        backward of the LAST micro-batch, whose kernels add to the buffer that already holds the earlier ones."""
        eng = self.engine
        prev_hook, prev_unit = eng.grad_ready_hook, eng.unit_loss_grad
        eng.grad_ready_hook, eng.unit_loss_grad = (self.bucketer.ready if sync else None), True
        try:
            loss.backward()
        finally:
            eng.grad_ready_hook, eng.unit_loss_grad = prev_hook, prev_unit
        if sync:
            self.bucketer.finish()

    def step(self, max_grad_norm: Optional[float] = 1.0, micro_batches: int = 1):
        self.optimizer.step(max_grad_norm=max_grad_norm, grad_scale=1.0 / (self.world * micro_batches))

    def train_step_device(self, x, max_grad_norm: Optional[float] = 1.0):
        """zero_grad -> get_loss -> backward (+ overlapped all-reduce) -> fused clip + AdamW. Device metrics.
        x: this rank's batch, or a list of micro-batches (gradient accumulation: BASELINE config 4's global batch of
        2048 at 2 / 4 GPUs is 4 / 2 micro-batches of 256 per rank; one all-reduce per optimiser step)."""
        self.model.train()
        xs = list(x) if isinstance(x, (list, tuple)) else [x]
        self.optimizer.zero_grad()
        metrics = None
        for i, xm in enumerate(xs):
            loss, m = self.get_loss(xm)
            self.backward(loss, sync=(i == len(xs) - 1))
            m = {k: v.detach() for k, v in m.items()}
            metrics = m if metrics is None else {k: metrics[k] + m[k] for k in m}
        if len(xs) > 1:
            metrics = {k: v / len(xs) for k, v in metrics.items()}
        self.step(max_grad_norm, micro_batches=len(xs))
        return metrics

    def reduce_metrics(self, scalars: torch.Tensor) -> torch.Tensor:
        """Mean over ranks of a small tensor of per-rank scalars (call at logging time, not every step)."""
        out = scalars.detach().clone()
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=self.group)
        return out / self.world
