"""Caches of tensors derived from parameters must notice every way the reference changes parameters.

The reference's default optimiser (general/radam.py:82,147) writes weights through ``p.data.copy_`` - invisible to autograd's
version counters - so the 3xTF32 weight split, the fused projection weights, the built 1x1-convolution matrices and captured
CUDA graphs are keyed on ``ops.param_epoch()`` as well, which every ``Optimizer.step`` advances.  CPU-only host logic here;
``tests/test_gpu_reference_training.py`` exercises the same on the device with the reference's own RAdam.
"""
import pytest
import torch

from categoricalnf_b200 import ops


class DataWritingSGD(torch.optim.Optimizer):
    """Updates like the reference's RAdam: through ``p.data`` (general/radam.py:82 ``p.data.copy_(p_data_fp32)``)."""

    def __init__(self, params, lr=0.1):
        super().__init__(params, dict(lr=lr))

    def step(self, closure=None):
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is not None:
                    p.data.copy_(p.data - group["lr"] * p.grad.data)


def test_data_writes_do_not_move_the_version_counter():
    p = torch.nn.Parameter(torch.randn(4, 4))
    v = p._version
    p.data.add_(1.0)
    assert p._version == v          # the premise of this file


def test_weight_split_follows_data_writing_optimizer():
    p = torch.nn.Parameter(torch.randn(8, 12))
    opt = DataWritingSGD([p])
    with torch.no_grad():
        hi0, lo0 = ops.weight_split(p)
        again = ops.weight_split(p)
        assert again[0] is hi0 and again[1] is lo0                  # cached while nothing changed
    p.grad = torch.ones_like(p)
    epoch = ops.param_epoch()
    opt.step()
    assert ops.param_epoch() > epoch                                 # global post-step hook
    with torch.no_grad():
        hi1, lo1 = ops.weight_split(p)
    want = p.detach()
    assert torch.equal(hi1, ops._rna_tf32(want))
    assert torch.allclose((hi1.double() + lo1.double()).float(), want, rtol=0, atol=1e-6)
    assert not torch.equal(hi1, hi0)


def test_weight_split_is_not_cached_while_gradients_are_recorded():
    p = torch.nn.Parameter(torch.randn(8, 12))
    a = ops.weight_split(p)                    # grad mode on, requires_grad: fresh
    p.data.mul_(2.0)                           # no optimizer, no version change, no epoch change
    b = ops.weight_split(p)
    assert torch.equal(b[0], ops._rna_tf32(p.detach())) and not torch.equal(a[0], b[0])
    with torch.no_grad():
        c = ops.weight_split(p, use_cache=False)
    assert torch.equal(c[0], b[0])


def test_invalidate_caches_and_mode_switch_advance_the_epoch():
    from categoricalnf_b200.layers.networks.linear import TCLinear
    e = ops.param_epoch()
    ops.invalidate_caches()
    assert ops.param_epoch() == e + 1
    lin = TCLinear(4, 4)
    e = ops.param_epoch()
    lin.train()                                # already training: nothing changes
    assert ops.param_epoch() == e
    lin.eval()
    assert ops.param_epoch() == e + 1
    fp = ops.param_fingerprint(list(lin.parameters()))
    with torch.no_grad():
        lin.weight.add_(1.0)                   # visible to the version counter
    assert ops.param_fingerprint(list(lin.parameters())) != fp


def test_invconv_eval_cache_key_follows_the_epoch():
    from categoricalnf_b200.layers.flows.permutation_layers import InvertibleConv
    conv = InvertibleConv(4).eval()
    k0 = conv._param_version()
    conv.l.data.add_(0.1)
    assert conv._param_version() == k0         # p.data write alone is invisible ...
    ops.invalidate_caches()
    assert conv._param_version() != k0         # ... the epoch is what invalidates


def test_chess_mask_is_the_reference_pattern():
    """coupling_layer.py:115-120: cat([ones(n,1), zeros(n,1)], dim=1).view(-1,1) alternates for even lengths."""
    from categoricalnf_b200.layers.flows.coupling_layer import CouplingLayer
    for n in (2, 4, 6, 10):
        half = n // 2
        want = torch.cat([torch.ones(n - half, 1), torch.zeros(half, 1)], dim=1).view(-1, 1)
        got = CouplingLayer.create_chess_mask(n)
        assert got.shape == (n, 1) and torch.equal(got, want)
    assert CouplingLayer.create_chess_mask(4).flatten().tolist() == [1.0, 0.0, 1.0, 0.0]
    with pytest.raises(RuntimeError):
        CouplingLayer.create_chess_mask(3)     # upstream fails on odd lengths too


def test_inplace_ldj_must_not_be_copied():
    with pytest.raises(RuntimeError):
        ops._ldj(torch.zeros(4), 4)            # CPU tensor: no fallback
