"""Autoregressive mixture coupling (reference layers/flows/autoregressive_coupling.py:13-54):
no mask - every channel is transformed from the output of an autoregressive network; forward only;
accumulates into the incoming ldj."""
import torch
import torch.nn as nn

from ... import functional as CF
from .flow_layer import FlowLayer


class AutoregressiveMixtureCDFCoupling(FlowLayer):

    def __init__(self, c_in, model_func, block_type=None, num_mixtures=10):
        super().__init__()
        self.c_in = c_in
        self.num_mixtures = num_mixtures
        self.block_type = block_type
        self.scaling_factor = nn.Parameter(torch.zeros(self.c_in))
        self.mixture_scaling_factor = nn.Parameter(torch.zeros(self.c_in, self.num_mixtures))
        self.nn = model_func(c_out=c_in * (2 + 3 * self.num_mixtures))

    def forward(self, z, ldj=None, reverse=False, **kwargs):
        if reverse:
            raise NotImplementedError
        if ldj is None:
            ldj = z.new_zeros(z.size(0))
        nn_out = self.nn(x=z, **kwargs)
        # the transform itself ignores the padding mask (upstream passes none, :38); only the
        # output is blanked afterwards (:44-45)
        z_out, layer_ldj, _ = CF.mixcdf(z, nn_out, self.num_mixtures, self.scaling_factor,
                                        self.mixture_scaling_factor)
        pad = kwargs.get("channel_padding_mask", None)
        if pad is not None:
            z_out = z_out * pad
        return z_out, ldj + layer_ldj

    def info(self):
        text = "Autoregressive Mixture CDF Coupling Layer - Input size %i" % self.c_in
        if self.block_type is not None:
            text += ", block type %s" % self.block_type
        return text
