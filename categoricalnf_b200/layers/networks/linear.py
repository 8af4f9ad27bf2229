"""Dense projections of the coupling networks on the tcgen05 tensor cores.

:class:`TCLinear` is ``nn.Linear`` (same constructor, parameter names ``weight`` / ``bias``, so reference
checkpoints load) whose forward runs ``cnf_linear_fwd``: TMA-fed ``tcgen05.mma`` with the accumulator in
tensor memory, 3xTF32 by default so that the flow transforms fed by the projection keep their 1e-4 parity.
:func:`convert_linears` swaps every ``nn.Linear`` of a coupling network (reference
layers/networks/graph_layers.py:24-25,64-71,192-202,307-315,402-405,574-577,712-716,766-779;
help_layers.py:57-124) for a :class:`TCLinear` sharing the same parameters.

The *final* projection of a coupling network can additionally be fused with the mixture transform
(``cnf_linear_mixcdf_fwd``): a network opts in by exposing ``cnf_features(x, **kw)`` (everything up to the
last projection) and ``cnf_final_linear`` (that ``nn.Linear``); an ``nn.Sequential`` ending in a Linear is
recognised as is.  See :func:`split_final_linear`.
"""
from __future__ import annotations

import torch
import torch.nn as nn

import os

from ... import ops

# Precision of the two backward GEMMs (grad_x, grad_W).  None = the forward's precision (3xTF32 by default: gradients
# at fp32 level, like the reference, which trains in plain fp32).  "tf32" (or CNF_B200_BWD_PRECISION=tf32) runs them as
# one-pass TF32 - a third of the tensor-core work, ~1e-3 relative gradient error.
BACKWARD_PRECISION = os.environ.get("CNF_B200_BWD_PRECISION") or None


class _TCLinearFn(torch.autograd.Function):
    """y = x W^T + b with all three GEMMs (forward, grad_x, grad_W) on the tcgen05 kernel.  The two backward
    products read their operands in place through MN-major shared-memory descriptors (``cnf_linear_bwd``)."""

    @staticmethod
    def forward(ctx, x, weight, bias, precision):
        ctx.precision = precision
        ctx.has_bias = bias is not None
        # grad mode is off inside Function.forward, so say explicitly that this is a training step: no cached split.  The
        # split made here (parameters / weight blocks only) is kept for the grad_x product of the backward pass.
        split = None
        if precision == "3xtf32" and weight.dtype == torch.float32 and weight.is_contiguous():
            split = ops.weight_split(weight, use_cache=False)
        ctx.has_split = split is not None
        ctx.save_for_backward(x, weight, *(split or ()))
        return ops.linear(x, weight, bias, precision=precision, cache_weight=False, split=split)

    @staticmethod
    def backward(ctx, gy):
        x, weight = ctx.saved_tensors[:2]
        split = tuple(ctx.saved_tensors[2:4]) if ctx.has_split else None
        gx, gw, gb = ops.linear_bwd(x, weight, gy, need_x=ctx.needs_input_grad[0], need_weight=ctx.needs_input_grad[1],
                                    need_bias=ctx.has_bias and ctx.needs_input_grad[2],
                                    precision=BACKWARD_PRECISION or ctx.precision, weight_split=split)
        return gx, gw, gb, None


class TCLinear(nn.Linear):
    """Drop-in ``nn.Linear`` evaluated by the tcgen05 kernel.  ``precision``: "3xtf32" (default) | "tf32"."""

    def __init__(self, in_features, out_features, bias=True, precision="3xtf32", device=None, dtype=None):
        super().__init__(in_features, out_features, bias=bias, device=device, dtype=dtype)
        self.precision = precision

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("categoricalnf_b200.TCLinear runs on CUDA only (input lives on %s)" % x.device)
        if torch.is_grad_enabled() and (x.requires_grad or self.weight.requires_grad):
            return _TCLinearFn.apply(x, self.weight, self.bias, self.precision)
        return ops.linear(x, self.weight, self.bias, precision=self.precision)

    def train(self, mode=True):
        if bool(mode) != self.training:      # train <-> eval switch: parameters may have moved through p.data
            ops.invalidate_caches()
        return super().train(mode)

    def extra_repr(self):
        return super().extra_repr() + ", tcgen05 %s" % self.precision

    @classmethod
    def from_linear(cls, lin: nn.Linear, precision="3xtf32"):
        """A TCLinear that SHARES ``lin``'s parameters (no copy)."""
        mod = cls.__new__(cls)
        nn.Module.__init__(mod)
        mod.in_features, mod.out_features = lin.in_features, lin.out_features
        mod.weight = lin.weight
        if lin.bias is None:
            mod.register_parameter("bias", None)
        else:
            mod.bias = lin.bias
        mod.precision = precision
        return mod


def convert_linears(module: nn.Module, precision="3xtf32", min_features=1) -> int:
    """Replace every ``nn.Linear`` below ``module`` by a parameter-sharing :class:`TCLinear`.
    Returns the number of layers converted; state-dict keys are unchanged."""
    n = 0
    for name, child in list(module.named_children()):
        if type(child) is nn.Linear and child.in_features >= min_features:
            setattr(module, name, TCLinear.from_linear(child, precision))
            n += 1
        else:
            n += convert_linears(child, precision, min_features)
    return n


def split_final_linear(net):
    """-> ``(features_fn, linear)`` when ``net`` = features followed by one last ``nn.Linear`` that the
    coupling layer may fuse with its transform, else ``None``.

    * opt-in protocol: ``net.cnf_final_linear`` (nn.Linear) and ``net.cnf_features(x, **kw)``;
    * ``nn.Sequential`` whose last module is an ``nn.Linear``.
    """
    lin = getattr(net, "cnf_final_linear", None)
    if isinstance(lin, nn.Linear) and callable(getattr(net, "cnf_features", None)):
        return net.cnf_features, lin
    if isinstance(net, nn.Sequential) and len(net) > 0 and isinstance(net[-1], nn.Linear):
        body = net[:-1]
        return (lambda x, **kw: body(x)), net[-1]
    return None
