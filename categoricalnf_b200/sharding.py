"""Batch sharding across the GPUs of one box: one process per GPU, ``torch.distributed`` (NCCL over
NVLink / NVSwitch on the GPU box, gloo in CPU tests) for the plumbing.

Samples are independent on the whole hot path (every ldj is per sample), so
the data path needs no collective: each rank runs the flow on a contiguous shard of the batch.
The only exchanges are
  * one all-reduce of (sum log-likelihood, sample count) per step            -> ``allreduce_loglik``
  * the per-channel (sum x, sum x^2, n) of the ActNorm data-dependent init   -> ``allreduce_moments``
  * gradients when training: flat buckets reduced on a communication stream
    while backward still runs                                                 -> ``GradientReducer`` (``allreduce_gradients``:
                                                                                 the simple post-backward form)
  * the log-likelihood pair off the compute stream, double-buffered          -> ``LogLikAllReducer``
This replaces the reference's single-process ``nn.DataParallel`` wrapper (general/mutils.py:243-249),
which re-broadcasts every parameter and gathers all outputs on GPU 0 each step.
"""
from __future__ import annotations

import math
from typing import Iterable, Optional, Tuple

import torch
import torch.distributed as dist


def world(group=None) -> Tuple[int, int]:
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_bounds(n: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) of ``n`` samples for ``rank``; the first ``n % world_size`` ranks get one
    extra sample, so ragged batches (and n < world_size: empty shards) are legal."""
    base, extra = divmod(n, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(t, rank: Optional[int] = None, world_size: Optional[int] = None, group=None):
    """This rank's slice (a view) along dim 0 of a tensor - or of every tensor in a tuple / list /
    dict; non-tensors pass through (e.g. scalars like ``beta``)."""
    if rank is None or world_size is None:
        rank, world_size = world(group)
    if isinstance(t, torch.Tensor):
        lo, hi = shard_bounds(t.shape[0], rank, world_size)
        return t[lo:hi]
    if isinstance(t, dict):
        return {k: shard_batch(v, rank, world_size) for k, v in t.items()}
    if isinstance(t, (tuple, list)):
        return type(t)(shard_batch(v, rank, world_size) for v in t)
    return t


def allreduce_loglik(ll_local: torch.Tensor, weight_local: Optional[torch.Tensor] = None, group=None):
    """Global (sum of log-likelihoods, sample count) from per-sample values of this rank's shard:
    ONE all-reduce of two float64 numbers.  ``weight_local`` (e.g. sequence lengths) replaces the
    count by a weighted one.  Returns a 2-element float64 tensor on the input's device."""
    acc = torch.empty(2, dtype=torch.float64, device=ll_local.device)
    acc[0] = ll_local.sum(dtype=torch.float64)
    acc[1] = float(ll_local.numel()) if weight_local is None else weight_local.sum(dtype=torch.float64)
    if world(group)[1] > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
    return acc


def bits_per_dim(acc: torch.Tensor, dims_per_sample: float = 1.0) -> float:
    """Mean negative log-likelihood in bits per dimension from ``allreduce_loglik``'s result
    (general/task.py:148-149)."""
    return float(-acc[0] / (acc[1] * dims_per_sample) * math.log2(math.e))


def allreduce_moments(x: torch.Tensor, pad: Optional[torch.Tensor] = None, group=None):
    """Per-channel (bias, scales) of the ActNorm data-dependent init (activation_normalization.py:
    55-67) over the *global* batch: local (sum x, sum x^2, n) in float64, one all-reduce, then
    bias = -mean and scales = -0.5 log var computed identically on every rank."""
    C = x.shape[-1]
    xd = x.reshape(-1, C).double()
    if pad is not None:
        w = pad.reshape(-1, 1).double()
        mom = torch.cat([(xd * w).sum(0), (xd * xd * w).sum(0), w.sum().expand(1)])
    else:
        mom = torch.cat([xd.sum(0), (xd * xd).sum(0), torch.tensor([float(xd.shape[0])], dtype=torch.float64,
                                                                 device=x.device)])
    if world(group)[1] > 1:
        dist.all_reduce(mom, op=dist.ReduceOp.SUM, group=group)
    n = mom[2 * C]
    mean = mom[:C] / n
    var = mom[C:2 * C] / n - mean * mean
    return (-mean).float(), (-0.5 * var.log()).float()


class LogLikAllReducer:
    """The per-step all-reduce of (sum log-likelihood, sample count) taken OFF the compute stream.

    ``ops.logistic_logprob(z, add=ldj, total=slot)`` leaves the rank's (sum, count) pair in ``slot`` from the epilogue of the
    kernel that finishes the log-likelihood; ``reduce(slot)`` records an event on the compute stream and issues the NCCL
    all-reduce of those 16 bytes on a dedicated communication stream, so the next step's kernels never queue behind the
    collective's launch latency (measured in round 1: ~0.1 ms per 2 ms step when issued in line).  ``slots`` rotating
    buffers keep up to that many steps in flight; ``result(i)`` waits for step i's reduction only."""

    def __init__(self, device, slots: int = 4, group=None):
        self.device, self.group = torch.device(device), group
        self.ws = world(group)[1]
        self.slots = [torch.zeros(2, dtype=torch.float64, device=self.device) for _ in range(slots)]
        self.done = [None] * slots
        self.step = 0
        self.stream = torch.cuda.Stream(device=self.device) if self.device.type == "cuda" else None

    def slot(self) -> torch.Tensor:
        """Buffer for the current step's (sum, count); waits (on the compute stream) until its previous reduction is done."""
        i = self.step % len(self.slots)
        if self.done[i] is not None and self.stream is not None:
            torch.cuda.current_stream(self.device).wait_event(self.done[i])
        return self.slots[i]

    def reduce(self) -> int:
        """All-reduce the current slot (filled by kernels already queued on the compute stream); returns the step index."""
        i = self.step % len(self.slots)
        if self.ws > 1:
            if self.stream is None:          # CPU tensors (gloo tests): in line
                dist.all_reduce(self.slots[i], op=dist.ReduceOp.SUM, group=self.group)
            else:
                ready = torch.cuda.Event()
                ready.record(torch.cuda.current_stream(self.device))
                with torch.cuda.stream(self.stream):
                    self.stream.wait_event(ready)
                    dist.all_reduce(self.slots[i], op=dist.ReduceOp.SUM, group=self.group)
                    self.done[i] = torch.cuda.Event()
                    self.done[i].record(self.stream)
        self.step += 1
        return self.step - 1

    def result(self, step: int) -> torch.Tensor:
        """Global (sum log-likelihood, count) of ``step`` as a CPU float64 tensor (blocks for that step only)."""
        i = step % len(self.slots)
        if self.stream is not None:
            if self.done[i] is not None:
                self.done[i].synchronize()
            else:
                torch.cuda.current_stream(self.device).synchronize()
        return self.slots[i].cpu()

    def finish(self) -> None:
        """Make the compute stream wait for every reduction still in flight (before timing ends / buffers are reused)."""
        if self.stream is not None:
            torch.cuda.current_stream(self.device).wait_stream(self.stream)


class GradientReducer:
    """Gradient all-reduce for one-process-per-GPU training, bucketed and OVERLAPPED with the backward pass
    (replaces nn.DataParallel's gather + re-broadcast, general/mutils.py:243-249).

    * Parameters are grouped into buckets in REVERSE registration order (roughly the order backward produces gradients);
      every bucket owns one pre-allocated flat fp32 buffer and each ``p.grad`` is a VIEW into it - no ``torch.cat`` staging,
      no copy back.
    * A ``register_post_accumulate_grad_hook`` per parameter counts the bucket down; when its last gradient has been
      accumulated the bucket's all-reduce is issued on a communication stream (after an event on the compute stream), so
      NCCL moves bucket k over NVLink while backward still computes bucket k+1.
    * ``finish()`` issues whatever is left (parameters that received no gradient contribute zeros), waits, and averages.
      Equal shards: each rank's loss is its local mean, ``finish()`` divides by the world size.  Ragged / empty shards
      (``shard_bounds``): ``loss = reducer.weight_loss(loss, n_local)`` before ``backward()`` and ``finish(n_local)`` give the
      exact global mean instead of a mean of means.
    * ``zero_grad()`` zeroes the flat buffers and keeps the views (``optimizer.zero_grad(set_to_none=True)`` would drop
      them; if a training loop does that anyway the hook copies the fresh gradient into its slot - still correct,
      one small copy per parameter slower)."""

    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_bytes: int = 32 << 20, group=None, profile=False,
                 hooks: bool = True):
        """``hooks=False``: no autograd hooks are registered - for gradients produced outside autograd's view (CUDA-graph
        replay of a whole training step); use :meth:`reduce_now` + :meth:`finish` after each step."""
        self.group = group
        self.profile = profile     # record CUDA events around every bucket's all-reduce on the communication stream
        self.comm_events = []
        self.ws = world(group)[1]
        self.params = [p for p in params if p.requires_grad]
        self.buckets = []          # dicts: flat, params, views, pending, handle
        self._owner = {}
        cur, size = [], 0
        for p in reversed(self.params):
            cur.append(p)
            size += p.numel() * 4
            if size >= bucket_bytes:
                self._make_bucket(cur)
                cur, size = [], 0
        if cur:
            self._make_bucket(cur)
        dev = self.params[0].device if self.params else torch.device("cpu")
        self.device = dev
        self.stream = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
        self.count = torch.zeros(1, dtype=torch.float64, device=dev)
        self.bytes_reduced = 0
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params] if hooks else []
        if hooks:
            self.zero_grad()

    def _make_bucket(self, ps):
        n = sum(p.numel() for p in ps)
        flat = torch.zeros(n, dtype=torch.float32, device=ps[0].device)
        views, off = [], 0
        for p in ps:
            views.append(flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        b = dict(flat=flat, params=list(ps), views=views, pending=len(ps), issued=False, done=None)
        for p, v in zip(ps, views):
            self._owner[id(p)] = (b, v)
        self.buckets.append(b)

    def zero_grad(self):
        for b in self.buckets:
            b["flat"].zero_()
            b["pending"], b["issued"], b["done"] = len(b["params"]), False, None
            for p, v in zip(b["params"], b["views"]):
                p.grad = v

    def _on_grad(self, p):
        b, v = self._owner[id(p)]
        if p.grad is not v:                       # the training loop replaced / dropped the view: put the gradient back
            v.copy_(p.grad)
            p.grad = v
        b["pending"] -= 1
        if b["pending"] == 0:
            self._issue(b)

    def _issue(self, b):
        if b["issued"]:
            return
        b["issued"] = True
        if self.ws == 1:
            return
        self.bytes_reduced += b["flat"].numel() * 4
        if self.stream is None:
            dist.all_reduce(b["flat"], op=dist.ReduceOp.SUM, group=self.group)
            return
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            if self.profile:
                e0 = torch.cuda.Event(enable_timing=True)
                e0.record(self.stream)
            dist.all_reduce(b["flat"], op=dist.ReduceOp.SUM, group=self.group)
            b["done"] = torch.cuda.Event(enable_timing=self.profile)
            b["done"].record(self.stream)
            if self.profile:
                self.comm_events.append((b["flat"].numel() * 4, e0, b["done"]))

    def reduce_now(self):
        """For gradients that did NOT come through the hooks - e.g. left in ``p.grad`` by a CUDA-graph replay of the whole
        training step (``GraphedTrainingStep``: no autograd runs at replay time) - copy them into the flat buckets (one
        multi-tensor copy per bucket), point ``p.grad`` at the views and issue every bucket's all-reduce on the
        communication stream.  Follow with :meth:`finish`."""
        for b in self.buckets:
            src, dst = [], []
            for p, v in zip(b["params"], b["views"]):
                if p.grad is None:
                    v.zero_()
                elif p.grad is not v and p.grad.data_ptr() != v.data_ptr():
                    src.append(p.grad)
                    dst.append(v)
            if src:
                torch._foreach_copy_(dst, src)
            for p, v in zip(b["params"], b["views"]):
                p.grad = v
            b["pending"], b["issued"] = 0, False
            self._issue(b)

    def comm_stats(self):
        """(bytes reduced, ms on the communication stream, bus GB/s = 2 (N-1) / N x bytes / time) of the recorded buckets
        (``profile=True``); clears the record.  Call after a synchronise."""
        nbytes = sum(n for n, _, _ in self.comm_events)
        ms = sum(a.elapsed_time(b) for _, a, b in self.comm_events)
        self.comm_events = []
        bus = (2.0 * (self.ws - 1) / self.ws) * nbytes / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        return nbytes, ms, bus

    def weight_loss(self, loss: torch.Tensor, local_count: int) -> torch.Tensor:
        """Ragged shards: turn this rank's local-MEAN loss into its local SUM (``loss * local_count``) BEFORE ``backward()``
        - buckets are reduced while backward runs, so the weight has to be in the gradients already.  Together with
        ``finish(local_count)`` the result is g = sum_r n_r g_r / sum_r n_r, the exact gradient of the global mean; an
        empty shard contributes weight 0 (not a zero that is still counted in the divisor)."""
        return loss * float(local_count)

    def finish(self, local_count: Optional[int] = None):
        """Complete the step's reduction: issue the buckets whose parameters received no gradient (their zeros are
        reduced so every rank issues the same collectives), wait for the communication stream, divide.
        ``local_count`` None: equal shards, every rank's loss is its local mean -> divide by the world size.
        ``local_count`` given: the loss went through ``weight_loss`` -> divide by the all-reduced global sample count."""
        for b in self.buckets:
            if not b["issued"]:
                self._issue(b)
        if self.ws == 1:
            if local_count is not None:
                for b in self.buckets:
                    b["flat"].div_(max(float(local_count), 1.0))
            return
        if self.stream is not None:
            torch.cuda.current_stream(self.device).wait_stream(self.stream)
        if local_count is None:
            for b in self.buckets:
                b["flat"].div_(float(self.ws))
            return
        self.count.fill_(float(local_count))
        dist.all_reduce(self.count, op=dist.ReduceOp.SUM, group=self.group)
        inv = (1.0 / self.count.clamp(min=1.0)).to(torch.float32)
        for b in self.buckets:
            b["flat"].mul_(inv)

    def close(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []


def allreduce_gradients(params: Iterable[torch.nn.Parameter], bucket_bytes: int = 64 << 20, group=None,
                        local_count: Optional[int] = None) -> None:
    """Post-backward (blocking) gradient all-reduce in flat buckets - the simple form; ``GradientReducer`` is the overlapped
    one.  Each rank's gradients are those of its local-MEAN loss.  ``local_count`` (samples in this rank's shard) weights
    the average: g = sum_r n_r g_r / sum_r n_r, the exact gradient of the global mean for ragged and empty shards; None
    assumes equal shards (g = sum_r g_r / world_size).  Parameters without a gradient contribute zeros so every rank
    issues the same collectives."""
    ws = world(group)[1]
    if ws == 1:
        return
    params = [p for p in params if p.requires_grad]
    if not params:
        return
    weight = 1.0 if local_count is None else float(local_count)
    count = torch.tensor([weight if local_count is not None else 1.0], dtype=torch.float64, device=params[0].device)
    dist.all_reduce(count, op=dist.ReduceOp.SUM, group=group)
    denom = float(count.item())
    if denom <= 0:
        raise RuntimeError("allreduce_gradients: no samples on any rank")
    bucket, size = [], 0

    def flush():
        nonlocal bucket, size
        if not bucket:
            return
        flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in bucket])
        if weight != 1.0:
            flat.mul_(weight)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(denom)
        off = 0
        for p in bucket:
            n = p.numel()
            if p.grad is None:
                p.grad = flat[off:off + n].view_as(p).clone()
            else:
                p.grad.copy_(flat[off:off + n].view_as(p))
            off += n
        bucket, size = [], 0

    for p in params:
        bucket.append(p)
        size += p.numel() * p.element_size()
        if size >= bucket_bytes:
            flush()
    flush()


def broadcast_parameters(module: torch.nn.Module, src: int = 0, group=None) -> None:
    """Make every rank start from rank ``src``'s parameters and buffers (after a seeded or
    data-dependent initialisation)."""
    if world(group)[1] == 1:
        return
    from . import ops
    with torch.no_grad():      # broadcast into the tensor itself (not ``.data``) so its version counter moves
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t, src=src, group=group)
    ops.invalidate_caches()
