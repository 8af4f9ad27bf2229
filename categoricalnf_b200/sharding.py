"""Batch sharding across the GPUs of one box: one process per GPU, ``torch.distributed`` (NCCL over
NVLink / NVSwitch on the GPU box, gloo in CPU tests) for the plumbing.

Samples are independent on the whole hot path (every ldj is per sample), so
the data path needs no collective: each rank runs the flow on a contiguous shard of the batch.
The only exchanges are
  * one all-reduce of (sum log-likelihood, sample count) per step            -> ``allreduce_loglik``
  * the per-channel (sum x, sum x^2, n) of the ActNorm data-dependent init   -> ``allreduce_moments``
  * gradients when training (flat buckets, summed, averaged)                 -> ``allreduce_gradients``
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


def allreduce_gradients(params: Iterable[torch.nn.Parameter], bucket_bytes: int = 64 << 20, group=None) -> None:
    """Average gradients across ranks in flat buckets (NVSwitch: size buckets for launch latency,
    not link count).  Parameters without a gradient contribute zeros so every rank issues the
    same collectives."""
    ws = world(group)[1]
    if ws == 1:
        return
    params = [p for p in params if p.requires_grad]
    bucket, size = [], 0

    def flush():
        nonlocal bucket, size
        if not bucket:
            return
        flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(ws)
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
