"""Affine coupling layer ("AffineCoupling"; reference layers/flows/coupling_layer.py:10-129).

Same constructor, buffers (``mask``), parameters (``scaling_factor``, ``nn.*``) and quirks as the
reference; the transform itself is one launch of ``cnf_affine_coupling`` (csrc/elementwise.cu)
instead of ~10 eager ops.
"""
import math

import torch
import torch.nn as nn

from ... import functional as CF
from ..networks.help_layers import run_sequential_with_mask
from ._masks import broadcast_mask, mask_lists
from .flow_layer import FlowLayer


class CouplingLayer(FlowLayer):

    def __init__(self, c_in, mask, model_func, block_type=None, c_out=-1, **kwargs):
        super().__init__()
        self.c_in = c_in
        self.c_out = c_out if c_out > 0 else 2 * c_in
        self.register_buffer("mask", mask)
        self.block_type = block_type
        self.scaling_factor = nn.Parameter(torch.zeros(c_in))
        self.nn = model_func(c_out=self.c_out)

    # -- black-box conditioner (coupling_layer.py:28-39) ------------------------------------------
    def run_network(self, x, length=None, **kwargs):
        if isinstance(self.nn, nn.Sequential):
            nn_out = run_sequential_with_mask(self.nn, x, length=length, **kwargs)
        else:
            nn_out = self.nn(x, length=length, **kwargs)
        pad = kwargs.get("channel_padding_mask", None)
        if pad is not None:
            nn_out = nn_out * pad
        return nn_out

    def _prepare_mask(self, mask, z):
        return broadcast_mask(self.mask, z)

    def forward(self, z, ldj=None, reverse=False, channel_padding_mask=None, **kwargs):
        # channel_padding_mask is bound here and therefore never reaches the network (App. B #3),
        # and the ldj of this layer is not pad-masked (App. B #4) - both as upstream.
        if ldj is None:
            ldj = z.new_zeros(z.size(0))
        nn_out = self.run_network(x=z * self._prepare_mask(self.mask, z), **kwargs)
        mask_c, mask_s = mask_lists(self, "mask", z.size(1))
        z_out, layer_ldj = CF.affine_coupling(z, nn_out, self.scaling_factor, mask_c=mask_c, mask_s=mask_s,
                                              reverse=reverse)
        return z_out, ldj + layer_ldj

    # -- static helpers other files call directly ---------------------------------------------------
    @staticmethod
    def get_coup_params(nn_out, mask, scaling_factor=None):
        """Bounded (s, t) tensors (coupling_layer.py:76-85); parameter plumbing in eager torch."""
        pair = nn_out.view(nn_out.shape[:-1] + (nn_out.shape[-1] // 2, 2))
        s, t = pair[..., 0], pair[..., 1]
        if scaling_factor is not None:
            bound = scaling_factor.exp().view(1, 1, -1)
            s = torch.tanh(s / bound.clamp(min=1.0)) * bound
        keep = 1 - mask
        return s * keep, t * keep

    @staticmethod
    def run_with_params(orig_z, s, t, reverse=False):
        """(z, ldj) for explicit, already bounded (s, t) (coupling_layer.py:88-98)."""
        return CF.affine_explicit(orig_z, s, t, reverse=reverse)

    @staticmethod
    def create_channel_mask(c_in, ratio=0.5, mask_floor=True):
        """[1, c_in]: the first floor/ceil(c_in*ratio) channels condition, the rest are transformed."""
        n_cond = int(math.floor(c_in * ratio)) if mask_floor else int(math.ceil(c_in * ratio))
        mask = torch.zeros(1, c_in)
        mask[0, :n_cond] = 1.0
        return mask

    @staticmethod
    def create_chess_mask(seq_len=2):
        """[seq_len, 1].  Upstream concatenates a ones and a zeros COLUMN and flattens the [n, 2] result row by row
        (coupling_layer.py:115-120), which alternates conditioning / transformed positions: [1,0] for 2, [1,0,1,0] for 4.
        Odd lengths fail upstream (the two columns differ in height) and fail here."""
        assert seq_len > 1
        if seq_len % 2 != 0:
            raise RuntimeError("create_chess_mask: odd seq_len %d (the reference's torch.cat of a [%d,1] and a [%d,1] "
                               "column along dim 1 fails as well)" % (seq_len, seq_len - seq_len // 2, seq_len // 2))
        return torch.stack([torch.ones(seq_len // 2), torch.zeros(seq_len // 2)], dim=1).view(-1, 1)

    def info(self):
        kind = "channel" if self.mask.size(0) == 1 else "chess"
        text = "Coupling Layer - Input size %i" % self.c_in
        if self.block_type is not None:
            text += ", block type %s" % self.block_type
        return text + ", mask ratio %.2f, %s mask" % ((1 - self.mask).mean().item(), kind)
