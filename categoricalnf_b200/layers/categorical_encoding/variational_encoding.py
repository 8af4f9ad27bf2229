"""Variational encoding of categorical variables: class-conditional flow as encoder q(z|x), learned
decoder as p(x|z) (reference layers/categorical_encoding/variational_encoding.py:18-188).

Upstream notes (App. B #7): only ``num_flows=0`` (one ExtActNorm) is usable there - with
couplings its ``_flow_forward`` unpacks two values from MixtureCDFCoupling's triple.  The same
composition is kept here; the triple is tolerated instead of crashing.  The reverse pass of the
reference reads an unregistered ``category_prior``; it is unused for decoding and dropped.
"""
import torch
import torch.nn as nn

from ... import ops
from ..flows.activation_normalization import ExtActNormFlow
from ..flows.coupling_layer import CouplingLayer
from ..flows.distributions import LogisticDistribution
from ..flows.flow_layer import FlowLayer
from ..flows.mixture_cdf_layer import MixtureCDFCoupling
from ..flows.permutation_layers import InvertibleConv
from ..networks.help_layers import SimpleLinearLayer
from .decoder import _param, create_decoder, create_embed_layer
from .linear_encoding import LinearCategoricalEncoding


class VariationalCategoricalEncoding(FlowLayer):

    def __init__(self, num_dimensions, flow_config, dataset_class=None, vocab=None, vocab_size=-1, use_decoder=False,
                 decoder_config=None, default_embed_layer_dims=64, category_prior=None, **kwargs):
        super().__init__()
        self.use_decoder = use_decoder
        self.dataset_class = dataset_class
        self.D = num_dimensions
        self.embed_layer, self.vocab_size = create_embed_layer(vocab, vocab_size, default_embed_layer_dims)
        self.num_categories = self.vocab_size
        self.prior_distribution = LogisticDistribution(mu=0.0, sigma=1.0)
        self.flow_layers = _create_flows(num_dims=num_dimensions, embed_dims=self.embed_layer.weight.shape[1],
                                         config=flow_config)
        self.decoder = create_decoder(num_categories=self.vocab_size, num_dims=self.D, config=decoder_config)

    def forward(self, z, ldj=None, reverse=False, beta=1, delta=0.0, channel_padding_mask=None, u_noise=None, **kwargs):
        batch_size, seq_length = z.size(0), z.size(1)
        z = z.reshape((batch_size * seq_length, 1) + z.shape[2:])
        if channel_padding_mask is not None:
            pad = channel_padding_mask.reshape(batch_size * seq_length, 1, -1)
        else:
            pad = torch.ones((batch_size * seq_length, 1, 1), dtype=torch.float32, device=z.device)
        ldj_loc = torch.zeros(z.size(0), dtype=torch.float32, device=z.device)
        detailed_ldj = {}
        if not reverse:
            z_categ = z
            shape = (batch_size * seq_length, 1, self.D)
            if u_noise is None:
                z_cont = self.prior_distribution.sample(shape=shape).to(z_categ.device)
            else:
                z_cont = ops.logistic_sample(shape, z.device, noise=u_noise, mu=self.prior_distribution.mu,
                                             sigma=self.prior_distribution.sigma, eps=self.prior_distribution.eps)
            init_log_p = self.prior_distribution.log_prob(z_cont).sum(dim=[1, 2])
            z_cont, ldj_forward = self._flow_forward(z_cont, z_categ, reverse=False)
            class_prob_log = self._decoder_forward(z_cont, z_categ)
            # [N,1] x [N] broadcasting as upstream (:72-73) is only meaningful for N == 1; the
            # per-token value is what is meant
            class_prob_log = class_prob_log.reshape(-1)
            ldj_loc = (beta * class_prob_log - (init_log_p - ldj_forward)) * pad.squeeze()
            z_out = z_cont * pad
            if self.training:
                detailed_ldj = LinearCategoricalEncoding._stats(z_out, class_prob_log, pad)
            z_out = z_out.reshape(batch_size, seq_length, -1)
        else:
            assert z.size(-1) == self.D, \
                "[!] ERROR in categorical decoding: Input must have %i latent dimensions but got %i" % (self.D, z.shape[-1])
            z_out = self._decoder_sample(z).reshape(batch_size, seq_length)
        ldj_loc = ldj_loc.reshape(batch_size, seq_length).sum(dim=-1)
        ldj = ldj_loc if ldj is None else ldj + ldj_loc
        return z_out, ldj, detailed_ldj

    def _flow_forward(self, z_cont, z_categ, reverse, **kwargs):
        ldj = torch.zeros(z_cont.size(0), dtype=torch.float32, device=z_cont.device)
        embed_features = self.embed_layer(z_categ)
        for flow in (self.flow_layers if not reverse else reversed(self.flow_layers)):
            res = flow(z_cont, ldj, ext_input=embed_features, reverse=reverse, **kwargs)
            if len(res) == 3:    # MixtureCDFCoupling returns only its own ldj (App. B #1)
                z_cont, ldj = res[0], ldj + res[1]
            else:
                z_cont, ldj = res
        return z_cont, ldj

    def _decoder_forward(self, z_cont, z_categ, **kwargs):
        return self.decoder(z_cont).gather(dim=-1, index=z_categ.view(-1, 1, 1)).squeeze(-1)

    def _decoder_sample(self, z_cont, **kwargs):
        return self.decoder(z_cont).argmax(dim=-1)

    def info(self):
        s = "Variational Encodings of categories, with %i dimensions and %i flows.\n" % (self.D, len(self.flow_layers))
        s += "-> Decoder network: %s\n" % self.decoder.info()
        return s + "\n".join("-> [%i] " % (i + 1) + flow.info() for i, flow in enumerate(self.flow_layers))


def _create_flows(num_dims, embed_dims, config):
    """One ExtActNorm (``num_flows=0``) or ``num_flows`` x [ExtActNorm, InvertibleConv,
    MixtureCDFCoupling] with ``config["model_func"]`` as coupling net (:149-186)."""
    num_flows = _param(config, "num_flows", 0)

    def actnorm():
        return ExtActNormFlow(c_in=num_dims, net=SimpleLinearLayer(c_in=embed_dims, c_out=2 * num_dims, data_init=True))

    if num_flows == 0:
        return nn.ModuleList([actnorm()])
    if config is None or config.get("model_func", None) is None:
        raise KeyError("[!] ERROR: flow_config[\"model_func\"] is required for a variational encoding with couplings")
    model_func, block_type = config["model_func"], _param(config, "block_type", None)
    num_mixtures = _param(config, "num_mixtures", 8)
    if num_dims > 1:
        base = CouplingLayer.create_channel_mask(c_in=num_dims)
        mask_of = lambda i: base
    else:
        base = CouplingLayer.create_chess_mask()
        mask_of = lambda i: base if i % 2 == 0 else 1 - base
    layers = []
    for i in range(num_flows):
        layers += [actnorm(), InvertibleConv(c_in=num_dims),
                   MixtureCDFCoupling(c_in=num_dims, mask=mask_of(i), block_type=block_type, model_func=model_func,
                                      num_mixtures=num_mixtures)]
    return nn.ModuleList(layers)
