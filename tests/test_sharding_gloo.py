"""Host-side sharding logic on CPU with the gloo backend, world_size 2 (the N>1 path of bench.py):
shards of a ragged batch reassemble to the whole, the (sum log-lik, count) all-reduce equals the
single-process value, data-init moments agree with the oracle's global statistics, gradient
averaging matches the mean of the per-rank gradients."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from categoricalnf_b200 import sharding as SH
from oracle import cnf_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world_size, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        g = torch.Generator().manual_seed(0)
        B, S, C = 7, 5, 4      # ragged: 7 samples over 2 ranks -> 4 + 3
        x = torch.randn(B, S, C, generator=g)
        ll = torch.randn(B, generator=g)
        pad = (torch.rand(B, S, 1, generator=g) > 0.3).float()
        lo, hi = SH.shard_bounds(B, rank, world_size)
        xs, lls, pads = SH.shard_batch((x, ll, pad))
        assert xs.shape[0] == hi - lo and torch.equal(xs, x[lo:hi])
        acc = SH.allreduce_loglik(lls)
        bias, scales = SH.allreduce_moments(xs, pads)
        lin = torch.nn.Linear(C, 3)
        torch.manual_seed(0)
        with torch.no_grad():
            for p in lin.parameters():
                p.copy_(torch.randn(p.shape, generator=torch.Generator().manual_seed(5)))
        SH.broadcast_parameters(lin)
        lin(xs.reshape(-1, C)).pow(2).sum().backward()
        local_grad = lin.weight.grad.clone()
        SH.allreduce_gradients(lin.parameters(), bucket_bytes=16)
        out[rank] = dict(acc=acc, bias=bias, scales=scales, grad=lin.weight.grad.clone(), local_grad=local_grad,
                         bounds=(lo, hi))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process():
    world_size, port = 2, _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world_size, port, out), nprocs=world_size, join=True)
        res = {k: v for k, v in out.items()}
    g = torch.Generator().manual_seed(0)
    B, S, C = 7, 5, 4
    x = torch.randn(B, S, C, generator=g)
    ll = torch.randn(B, generator=g)
    pad = (torch.rand(B, S, 1, generator=g) > 0.3).float()
    assert res[0]["bounds"] == (0, 4) and res[1]["bounds"] == (4, 7)
    for r in range(world_size):
        assert torch.allclose(res[r]["acc"], torch.tensor([ll.double().sum(), float(B)], dtype=torch.float64))
    bias_ref, scales_ref = O.actnorm_data_init(x, pad)
    for r in range(world_size):
        assert torch.allclose(res[r]["bias"], bias_ref.flatten(), atol=1e-6)
        assert torch.allclose(res[r]["scales"], scales_ref.flatten(), atol=1e-5)
    mean_grad = (res[0]["local_grad"] + res[1]["local_grad"]) / 2
    for r in range(world_size):
        assert torch.allclose(res[r]["grad"], mean_grad, atol=1e-6)


def test_shard_bounds_cover_and_handle_tiny_batches():
    for n in (0, 1, 3, 8, 4097):
        for ws in (1, 2, 4, 8):
            spans = [SH.shard_bounds(n, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_single_process_is_identity():
    ll = torch.arange(6, dtype=torch.float32)
    acc = SH.allreduce_loglik(ll)
    assert acc.tolist() == [15.0, 6.0]
    assert abs(SH.bits_per_dim(acc, 2.0) - (-15.0 / 12.0 * 1.4426950408889634)) < 1e-12


# ---- overlapped gradient reducer / weighted averaging / log-likelihood reducer (world size 2, gloo) --------------------------
def _model():
    torch.manual_seed(3)
    return torch.nn.Sequential(torch.nn.Linear(4, 8), torch.nn.Tanh(), torch.nn.Linear(8, 8), torch.nn.Tanh(), torch.nn.Linear(8, 1))


def _reducer_worker(rank, world_size, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        g = torch.Generator().manual_seed(1)
        B = 7                                              # ragged: 4 + 3
        x, y = torch.randn(B, 4, generator=g), torch.randn(B, 1, generator=g)
        lo, hi = SH.shard_bounds(B, rank, world_size)
        model = _model()
        red = SH.GradientReducer(model.parameters(), bucket_bytes=64)       # several buckets
        assert len(red.buckets) >= 3
        res = {}
        for step in range(2):                              # second step: views survive zero_grad()
            red.zero_grad()
            loss = ((model(x[lo:hi]) - y[lo:hi]) ** 2).mean()
            red.weight_loss(loss, hi - lo).backward()
            assert all(b["issued"] for b in red.buckets)   # every bucket went out from the hooks, during backward
            red.finish(hi - lo)
            res["grads%d" % step] = [p.grad.clone() for p in model.parameters()]
            assert all(p.grad.data_ptr() == red._owner[id(p)][1].data_ptr() for p in model.parameters())
        # a loop that drops the views (optimizer.zero_grad(set_to_none=True)) still reduces correctly
        for p in model.parameters():
            p.grad = None
        for b in red.buckets:
            b["flat"].zero_()
            b["pending"], b["issued"] = len(b["params"]), False
        loss = ((model(x[lo:hi]) - y[lo:hi]) ** 2).mean()
        red.weight_loss(loss, hi - lo).backward()
        red.finish(hi - lo)
        res["grads_dropped"] = [p.grad.clone() for p in model.parameters()]
        red.close()
        # simple post-backward form with weights; an EMPTY shard on rank 1
        model2 = _model()
        n_local = B if rank == 0 else 0
        model2.zero_grad()
        if n_local:
            ((model2(x) - y) ** 2).mean().backward()
        SH.allreduce_gradients(model2.parameters(), bucket_bytes=64, local_count=n_local)
        res["grads_empty"] = [p.grad.clone() for p in model2.parameters()]
        # log-likelihood pair: rotating slots, reduce, result
        ll = torch.arange(B, dtype=torch.float64)[lo:hi]
        llr = SH.LogLikAllReducer("cpu", slots=2)
        steps = []
        for k in range(3):
            slot = llr.slot()
            slot[0], slot[1] = float(ll.sum()) + k, float(hi - lo)
            steps.append(llr.reduce())
            res["ll%d" % k] = llr.result(steps[-1]).clone()
        out[rank] = res
    finally:
        dist.destroy_process_group()


def test_overlapped_gradient_reducer_and_weighted_average():
    world_size, port = 2, _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_reducer_worker, args=(world_size, port, out), nprocs=world_size, join=True)
        res = {k: v for k, v in out.items()}
    g = torch.Generator().manual_seed(1)
    B = 7
    x, y = torch.randn(B, 4, generator=g), torch.randn(B, 1, generator=g)
    model = _model()
    ((model(x) - y) ** 2).mean().backward()                # single-process gradient of the GLOBAL mean
    want = [p.grad for p in model.parameters()]
    for r in range(world_size):
        for key in ("grads0", "grads1", "grads_dropped", "grads_empty"):
            for got, w in zip(res[r][key], want):
                assert torch.allclose(got, w, atol=1e-6), key
        for k in range(3):
            assert res[r]["ll%d" % k].tolist() == [float(sum(range(B))) + 2 * k, float(B)]


def test_reduce_now_without_hooks_single_process():
    """``GradientReducer(hooks=False)`` + ``reduce_now()``: gradients produced outside autograd's view (a CUDA-graph replay
    leaves them in ``p.grad``) are copied into the flat buckets, ``p.grad`` becomes the bucket view, values unchanged."""
    model = _model()
    red = SH.GradientReducer(model.parameters(), bucket_bytes=64, hooks=False)
    x = torch.randn(5, 4)
    model(x).sum().backward()
    want = [p.grad.clone() for p in model.parameters()]
    red.reduce_now()
    red.finish()
    for p, w in zip(model.parameters(), want):
        assert torch.equal(p.grad, w)
        assert p.grad.data_ptr() == red._owner[id(p)][1].data_ptr()
    red.close()


def _step_count_worker(rank, world_size, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        import bench
        # every rank derives a different count from its own timing (here: 500 + 37 rank) ...
        n = bench.same_on_all_ranks(500 + 37 * rank, torch.device("cpu"), dist)
        # ... and a loop of that many collectives only terminates when the counts agree
        t = torch.zeros(1)
        for _ in range(n % 7 + 1):
            dist.all_reduce(t)
        out[rank] = n
    finally:
        dist.destroy_process_group()


def test_bench_step_counts_are_agreed_across_ranks():
    """bench.py's sustained legs size their loops from a LOCAL timing; the count must be made identical on all ranks
    before it drives per-step all-reduces (a rank-local count hung the N = 2 run once): every rank takes the maximum."""
    world_size, port = 2, _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_step_count_worker, args=(world_size, port, out), nprocs=world_size, join=True)
        res = {k: v for k, v in out.items()}
    assert res[0] == res[1] == 537
    import bench
    assert bench.same_on_all_ranks(12, torch.device("cpu"), None) == 12
