"""Small conditioner networks the encodings construct themselves
(reference layers/networks/help_layers.py:57-124).  Dense stacks of ``TCLinear`` - ``nn.Linear`` with the reference's
parameter names (App. A: ``pred_net.layer.*``, ``layers.inp_layer.0.*``, ``layers.main_net.<i>.*``), evaluated by the
tcgen05 projection kernel (``cnf_linear_fwd`` / ``cnf_linear_bwd``, 3xTF32) like every other projection of the hot path; at
evaluation time a Linear followed by a GELU is ONE launch (the kernel's GELU epilogue).
"""
import math

import torch
import torch.nn as nn

from ... import ops
from .linear import TCLinear


def run_mlp(seq, x):
    """``seq(x)`` for an ``nn.Sequential`` of TCLinear / GELU modules.  Without autograd a ``TCLinear -> GELU`` pair runs as
    one kernel launch (GELU in the projection's epilogue); otherwise module by module (the backward kernels differentiate
    the projection, torch the activation)."""
    mods = list(seq)
    if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in seq.parameters())):
        return seq(x)
    i = 0
    while i < len(mods):
        m = mods[i]
        if isinstance(m, TCLinear) and i + 1 < len(mods) and isinstance(mods[i + 1], nn.GELU) and \
                getattr(mods[i + 1], "approximate", "none") == "none":
            x = ops.linear(x, m.weight, m.bias, precision=m.precision, activation="gelu")
            i += 2
        else:
            x = m(x)
            i += 1
    return x


class SimpleLinearLayer(nn.Module):
    """One Linear layer; with ``data_init`` the scale half starts at zero and the bias half is
    rescaled so that class means start well separated (help_layers.py:59-66)."""

    def __init__(self, c_in, c_out, data_init=False):
        super().__init__()
        self.layer = TCLinear(c_in, c_out)
        if data_init:
            half = c_out // 2
            with torch.no_grad():
                self.layer.weight[half:, :] = 0
                self.layer.weight.mul_(4.0 / math.sqrt(c_out / 2))
                self.layer.bias.zero_()

    def forward(self, x, **kwargs):
        return self.layer(x)

    def initialize_zeros(self):
        with torch.no_grad():
            self.layer.weight.zero_()
            self.layer.bias.zero_()


class LinearNet(nn.Module):
    """GELU MLP with an optional external input concatenated after the first layer
    (help_layers.py:76-107)."""

    def __init__(self, c_in, c_out, num_layers, hidden_size, ext_input_dims=0, zero_init=False):
        super().__init__()
        self.inp_layer = nn.Sequential(TCLinear(c_in, hidden_size), nn.GELU())
        blocks = []
        for i in range(num_layers):
            width_in = hidden_size + ext_input_dims if i == 0 else hidden_size
            blocks += [TCLinear(width_in, hidden_size), nn.GELU()]
        blocks.append(TCLinear(hidden_size, c_out))
        self.main_net = nn.Sequential(*blocks)
        if zero_init:
            with torch.no_grad():
                self.main_net[-1].weight.zero_()
                self.main_net[-1].bias.zero_()

    def forward(self, x, ext_input=None, **kwargs):
        h = run_mlp(self.inp_layer, x)
        if ext_input is not None:
            h = torch.cat([h, ext_input], dim=-1)
        return run_mlp(self.main_net, h)

    def set_bias(self, bias):
        # upstream assigns the tensor as is (help_layers.py:106-107); a float64 prior would then
        # break F.linear on torch >= 1.6 (App. B #9), so the dtype/device of the layer is kept.
        last = self.main_net[-1].bias
        last.data = bias.to(device=last.device, dtype=last.dtype)


def run_sequential_with_mask(net, x, length=None, channel_padding_mask=None, src_key_padding_mask=None,
                             length_one_hot=None, time_embed=None, gt=None, importance_weight=1,
                             detail_out=False, **kwargs):
    """Run an ``nn.Sequential`` conditioner, blanking padded positions on the way in and out
    (help_layers.py:111-124)."""
    if channel_padding_mask is None:
        out = net(x)
    else:
        out = net(x * channel_padding_mask) * channel_padding_mask
    return (out, dict()) if detail_out else out
