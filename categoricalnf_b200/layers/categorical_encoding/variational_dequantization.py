"""Variational dequantization (reference layers/categorical_encoding/variational_dequantization.py:15-98).

Encodes a discrete value ``x`` as ``x + u`` with ``u in [0,1]`` drawn from a conditional flow: uniform noise -> logit ->
``num_flows`` x [ActNorm(c_in=1), affine CouplingLayer(c_in=1, chess mask, net conditioned on the embedding of x)] ->
sigmoid.  Same constructor, sub-module names (``embed_layer``, ``flow_layers.<i>``, ``sigmoid_flow``) and return values as
the reference.  Kernels: ``cnf_sigmoid_flow`` for both ends (the final one also adds the discrete value, :47),
``cnf_actnorm`` / ``cnf_affine_coupling`` in between, ``cnf_dequant_floor`` for the inverse; the reference's three
host-synchronising asserts become bits of the device status word.

``u_noise`` (optional, z's shape, U(0,1)) replaces the internal ``torch.rand_like`` draw so a run can be replayed
against the reference / oracle.
"""
import torch
import torch.nn as nn

from ... import functional as CF
from ... import ops
from ..flows.activation_normalization import ActNormFlow
from ..flows.coupling_layer import CouplingLayer
from ..flows.flow_layer import FlowLayer
from ..flows.sigmoid_layer import ALPHA, SigmoidFlow
from .decoder import create_embed_layer


def _get(config, key, default=None, required=False):
    if config is None or key not in config or config[key] is None:
        if required:
            raise KeyError("flow_config[\"%s\"] is required for variational dequantization" % key)
        return default
    return config[key]


class VariationalDequantization(FlowLayer):

    def __init__(self, flow_config, vocab=None, vocab_size=-1, default_embed_layer_dims=128, **kwargs):
        super().__init__()
        self.embed_layer, self.vocab_size = create_embed_layer(vocab, vocab_size, default_embed_layer_dims)
        self.flow_layers = _create_flows(flow_config, self.embed_layer.weight.shape[1])
        self.sigmoid_flow = SigmoidFlow(reverse=True)

    def forward(self, z, ldj=None, reverse=False, u_noise=None, **kwargs):
        if ldj is None:
            ldj = z.new_zeros(z.size(0), dtype=torch.float32)
        if reverse:
            # the next lower whole number of every continuous value (:55-56)
            return ops.dequant_floor(z, self.vocab_size), ldj
        if u_noise is None:
            u_noise = torch.rand(z.shape, dtype=torch.float32, device=z.device)
        rand_inp = u_noise.to(torch.float32).reshape(z.shape).unsqueeze(dim=-1)
        rand_inp, ldj = self.sigmoid_flow(rand_inp, ldj=ldj, reverse=False)          # [0,1] -> R   (:40)
        rand_inp, ldj = self._flow_forward(rand_inp, z, ldj, **kwargs)               #  R    -> R   (:41)
        # R -> [0,1] and z.float() + noise in one launch (:42, :47); sigmoid keeps the noise inside [0,1] by construction
        z_out, ldj = CF.sigmoid_flow(rand_inp, ldj, reverse=False, alpha=ALPHA, add_tokens=z.unsqueeze(dim=-1))
        return z_out, ldj

    def _flow_forward(self, rand_inp, z, ldj, **kwargs):
        embed_features = self.embed_layer(z)
        for flow in self.flow_layers:
            rand_inp, ldj = flow(rand_inp, ldj, ext_input=embed_features, reverse=False, **kwargs)
        return rand_inp, ldj

    def info(self):
        s = "Variational Dequantization with %i flows.\n" % (len(self.flow_layers))
        s += "\n".join(["-> [%i] " % (i + 1) + flow.info() for i, flow in enumerate(self.flow_layers)])
        return s


def _create_flows(config, embed_dims):
    num_flows = _get(config, "num_flows", 4)
    model_func = _get(config, "model_func", required=True)
    block_type = _get(config, "block_type", None)
    layers = []
    for flow_index in range(num_flows):
        # c_in = 1: a 1x1 convolution would be a scalar, so a block is ActNorm + chess-mask coupling (:81-93)
        mask = CouplingLayer.create_chess_mask()
        if flow_index % 2 == 0:
            mask = 1 - mask
        layers += [ActNormFlow(c_in=1, data_init=False),
                   CouplingLayer(c_in=1, mask=mask, model_func=model_func, block_type=block_type)]
    return nn.ModuleList(layers)
