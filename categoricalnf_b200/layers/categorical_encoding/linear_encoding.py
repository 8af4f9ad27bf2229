"""Mixture-model / linear-flow encoding of categorical variables
(reference layers/categorical_encoding/linear_encoding.py:17-250).

Same constructor, sub-module names (``embed_layer``, ``flow_layers``, ``decoder``), buffer
(``category_prior``) and return triple as the reference.  For the mixture model (``num_flows=0``:
one class-conditional logistic per category) the whole forward - noise, logistic sample, class
affine, exact posterior over all V classes, per-sample ldj - is ONE kernel (``cnf_categ_encode``)
and decoding is one more (``cnf_categ_decode``); the reference materialises ``[B*S*V, 1, D]``
tensors through ~60 launches.  Linear flows (``num_flows>0``) and the decoder variant compose
the drop-in flow layers exactly as upstream does.
"""
import numpy as np
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import functional as CF
from ... import ops
from ..flows.activation_normalization import ExtActNormFlow
from ..flows.coupling_layer import CouplingLayer
from ..flows.distributions import LogisticDistribution
from ..flows.flow_layer import FlowLayer
from ..flows.permutation_layers import InvertibleConv
from ..networks.help_layers import LinearNet, SimpleLinearLayer
from .decoder import _param, create_decoder, create_embed_layer


def _host_noise() -> bool:
    """``CNF_B200_HOST_NOISE=1``: draw the encoding's uniform noise from torch's CPU generator with the reference's own call
    (``Uniform(0,1).sample((B*S, 1, D))`` on the host, then ``.to(device)`` - linear_encoding.py:78, distributions.py:139)
    instead of Philox inside the kernel.  Slower (one H2D copy per forward) but bit-identical noise to a reference run with
    the same seed - what ``tools/run_set_modeling.py`` uses to compare training runs step by step."""
    return os.environ.get("CNF_B200_HOST_NOISE", "0") not in ("", "0")


def philox_stream(device, n):
    """(seed, offset) for ``n`` in-kernel Philox draws, consumed from torch's CUDA generator so that
    ``torch.manual_seed`` makes the kernels reproducible.  Host-side bookkeeping only."""
    gen = torch.cuda.default_generators[device.index if device.index is not None else torch.cuda.current_device()]
    seed, offset = gen.initial_seed(), gen.get_offset()
    gen.set_offset(offset + 4 * ((n + 3) // 4))
    return seed & 0xFFFFFFFFFFFFFFFF, offset


class LinearCategoricalEncoding(FlowLayer):

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
        if self.use_decoder:
            self.decoder = create_decoder(num_categories=self.vocab_size, num_dims=self.D, config=decoder_config)
        if category_prior is None:
            category_prior = torch.zeros(self.vocab_size, dtype=torch.float32)
        else:
            assert category_prior.shape[0] == self.num_categories, \
                "[!] ERROR: Category prior needs to be of size [%i] but is %s" % (self.num_categories, str(category_prior.shape))
            if isinstance(category_prior, np.ndarray):
                category_prior = torch.from_numpy(category_prior)
        self.register_buffer("category_prior", F.log_softmax(category_prior.float(), dim=-1))

    # -- mixture model fast path ---------------------------------------------------------------------
    def _is_mixture_model(self):
        if self.use_decoder or len(self.flow_layers) != 1:
            return False
        flow = self.flow_layers[0]
        return isinstance(flow, ExtActNormFlow) and not flow.make_unique

    def class_table(self):
        """``[V, 2D]`` rows ``pred_net(embed(v))`` = (bias_v | raw log-scale_v): the only thing the
        class-conditional logistics depend on (:138; activation_normalization.py:127-128)."""
        return self.flow_layers[0].pred_net(self.embed_layer.weight)

    def _fused_ok(self, z):
        needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        return self._is_mixture_model() and not needs_grad and z.is_cuda

    def forward(self, z, ldj=None, reverse=False, beta=1, delta=0.0, channel_padding_mask=None, u_noise=None,
                **kwargs):
        """Encode ``z`` [B,S] int64 -> (z_cont [B,S,D], ldj [B], stats) or, with ``reverse``, decode
        ``z`` [B,S,D] -> tokens [B,S].  ``u_noise`` (optional, U(0,1) of B*S*D values) replaces the
        internal random draw - the hook parity tests use to replay the reference's noise."""
        batch_size, seq_length = z.size(0), z.size(1)
        detailed_ldj = {}
        if u_noise is None and not reverse and _host_noise():
            u_noise = torch.rand(batch_size * seq_length, 1, self.D).to(z.device)
        if self._fused_ok(z):
            with torch.no_grad():
                table = self.class_table()
            if not reverse:
                ldj_loc = torch.zeros(batch_size, dtype=torch.float32, device=z.device)
                seed, offset = (0, 0) if u_noise is not None else philox_stream(z.device, batch_size * seq_length * self.D)
                z_out, ldj_loc, cpl = ops.categ_encode(z.reshape(batch_size, seq_length), table, self.category_prior,
                                                       ldj_loc, noise=u_noise, seed=seed, offset=offset,
                                                       pad=channel_padding_mask, beta=float(beta),
                                                       want_class_prob=self.training)
                if self.training:
                    detailed_ldj = self._stats(z_out, cpl.reshape(-1), channel_padding_mask)
            else:
                assert z.size(-1) == self.D, \
                    "[!] ERROR in categorical decoding: Input must have %i latent dimensions but got %i" % (self.D, z.shape[-1])
                z_out = ops.categ_decode(z, table, self.category_prior)
                ldj_loc = torch.zeros(batch_size, dtype=torch.float32, device=z.device)
            ldj = ldj_loc if ldj is None else ldj + ldj_loc
            return z_out, ldj, detailed_ldj
        if self._is_mixture_model() and z.is_cuda and not reverse:
            # training: same kernel, differentiable in the class table (cnf_categ_encode_bwd); embed / pred_net receive
            # their gradients through table = pred_net(embed.weight), a [V,E] x [E,2D] product
            table = self.class_table()
            seed, offset = (0, 0) if u_noise is not None else philox_stream(z.device, batch_size * seq_length * self.D)
            z_out, ldj_loc, cpl = CF.categ_encode(z.reshape(batch_size, seq_length), table, self.category_prior, noise=u_noise,
                                                  seed=seed, offset=offset, pad=channel_padding_mask, beta=float(beta))
            if self.training:
                detailed_ldj = self._stats(z_out, cpl.reshape(-1), channel_padding_mask)
            ldj = ldj_loc if ldj is None else ldj + ldj_loc
            return z_out, ldj, detailed_ldj
        return self._composed_forward(z, ldj, reverse, beta, channel_padding_mask, u_noise, **kwargs)

    def try_forward_fused(self, z, actnorm, conv, beta=1, delta=0.0, channel_padding_mask=None, u_noise=None,
                          length=None, **kwargs):
        """Encode AND apply the first flow block's ``ActNormFlow`` + ``InvertibleConv`` in one kernel
        (evaluation only); None when not available for this configuration."""
        if self.training or not self._fused_ok(z):
            return None
        B, S = z.size(0), z.size(1)
        if not ops.categ_encode_fusable(B, S, self.num_categories, self.D):
            return None
        from ..flows.mixture_cdf_layer import add_next_block_ldj
        with torch.no_grad():
            table = self.class_table()
        weight, sldj = conv._get_weight(device_name=str(z.device), inverse=False)
        ldj = torch.zeros(B, dtype=torch.float32, device=z.device)
        seed, offset = (0, 0) if u_noise is not None else philox_stream(z.device, B * S * self.D)
        z_out, ldj, _ = ops.categ_encode(z.reshape(B, S), table, self.category_prior, ldj, noise=u_noise, seed=seed,
                                         offset=offset, pad=channel_padding_mask, beta=float(beta),
                                         fuse_next=(actnorm.bias, actnorm.scales, weight))
        add_next_block_ldj(ldj, actnorm, sldj, S, channel_padding_mask, length)
        return z_out, ldj, {}

    # -- general path: linear flows / decoder / training (composition of the flow layers) -----------
    def _composed_forward(self, z, ldj, reverse, beta, channel_padding_mask, u_noise, **kwargs):
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
            if not self.use_decoder:
                class_prior_log = torch.take(self.category_prior, z_categ.squeeze(dim=-1))
                log_point_prob = init_log_p - ldj_forward + class_prior_log
                class_prob_log = self._calculate_true_posterior(z_cont, z_categ, log_point_prob)
            else:
                class_prob_log = self._decoder_forward(z_cont, z_categ)
            ldj_loc = (beta * class_prob_log - (init_log_p - ldj_forward)) * pad.squeeze()
            z_out = z_cont * pad
            if self.training:
                detailed_ldj = self._stats(z_out, class_prob_log, pad)
            z_out = z_out.reshape(batch_size, seq_length, -1)
        else:
            assert z.size(-1) == self.D, \
                "[!] ERROR in categorical decoding: Input must have %i latent dimensions but got %i" % (self.D, z.shape[-1])
            z_out = self._posterior_sample(z) if not self.use_decoder else self._decoder_sample(z)
            z_out = z_out.reshape(batch_size, seq_length)
        ldj_loc = ldj_loc.reshape(batch_size, seq_length).sum(dim=-1)
        ldj = ldj_loc if ldj is None else ldj + ldj_loc
        return z_out, ldj, detailed_ldj

    @staticmethod
    def _stats(z_out, class_prob_log, pad):
        """Monitoring values returned as the third output while training (:95-107)."""
        with torch.no_grad():
            w = torch.ones_like(class_prob_log) if pad is None else pad.reshape(-1).float()
            n = w.sum()
            stats = {"avg_token_prob": (class_prob_log.exp() * w).sum() / n,
                     "avg_token_bpd": -(class_prob_log * w).sum() / n * np.log2(np.exp(1)),
                     "z_min": z_out.min(), "z_max": z_out.max(),
                     "z_std": z_out.reshape(-1, z_out.shape[-1]).std(0).mean()}
            return {k: v.detach() for k, v in stats.items()}

    def _flow_forward(self, z_cont, z_categ, reverse, **kwargs):
        ldj = torch.zeros(z_cont.size(0), dtype=torch.float32, device=z_cont.device)
        embed_features = self.embed_layer(z_categ)
        for flow in (self.flow_layers if not reverse else reversed(self.flow_layers)):
            z_cont, ldj = flow(z_cont, ldj, ext_input=embed_features, reverse=reverse, **kwargs)
        return z_cont, ldj

    def _decoder_forward(self, z_cont, z_categ, **kwargs):
        return self.decoder(z_cont).gather(dim=-1, index=z_categ.view(-1, 1))

    def _all_class_log_prob(self, z_cont, **kwargs):
        """log p(z | v) + log p(v) for every class v: inverse pass of all class-conditional flows (:155-164)."""
        n, V = z_cont.size(0), self.num_categories
        z_back_in = z_cont.expand(-1, V, -1).reshape(-1, 1, z_cont.size(2))
        sample_categ = torch.arange(V, dtype=torch.long, device=z_cont.device)[None, :].expand(n, -1).reshape(-1, 1)
        z_back, ldj_backward = self._flow_forward(z_back_in, sample_categ, reverse=True, **kwargs)
        back_log_p = self.prior_distribution.log_prob(z_back).sum(dim=[1, 2])
        return (back_log_p + ldj_backward).view(n, V) + self.category_prior[None, :]

    def _calculate_true_posterior(self, z_cont, z_categ, log_point_prob, **kwargs):
        denom = self._all_class_log_prob(z_cont, **kwargs)
        # the true class uses its forward-pass value (stability, :165-168)
        own = F.one_hot(z_categ.reshape(-1), num_classes=denom.size(1)).to(denom.dtype)
        denom = denom * (1 - own) + log_point_prob.unsqueeze(dim=-1) * own
        return log_point_prob - torch.logsumexp(denom, dim=-1)

    def _decoder_sample(self, z_cont, **kwargs):
        return self.decoder(z_cont).argmax(dim=-1)

    def _posterior_sample(self, z_cont, **kwargs):
        return self._all_class_log_prob(z_cont, **kwargs).argmax(dim=-1)

    def info(self):
        if len(self.flow_layers) > 1:
            s = "Linear Encodings of categories, with %i dimensions and %i flows.\n" % (self.D, len(self.flow_layers))
        else:
            s = "Mixture model encoding of categories with %i dimensions\n" % self.D
        s += "-> Prior distribution: %s\n" % self.prior_distribution.info()
        if self.use_decoder:
            s += "-> Decoder network: %s\n" % self.decoder.info()
        return s + "\n".join("-> [%i] " % (i + 1) + flow.info() for i, flow in enumerate(self.flow_layers))


def _create_flows(num_dims, embed_dims, config):
    """``num_flows`` x [ExtActNorm, InvertibleConv, affine CouplingLayer(LinearNet)], or a single
    ExtActNorm for the mixture model (:214-250)."""
    num_flows = _param(config, "num_flows", 0)
    hidden_layers = _param(config, "hidden_layers", 2)
    hidden_size = _param(config, "hidden_size", 256)

    def coupling_net(c_out):
        return LinearNet(c_in=num_dims, c_out=c_out, num_layers=hidden_layers, hidden_size=hidden_size,
                         ext_input_dims=embed_dims)

    def actnorm():
        return ExtActNormFlow(c_in=num_dims, net=SimpleLinearLayer(c_in=embed_dims, c_out=2 * num_dims, data_init=True))

    if num_flows == 0 or num_dims == 1:
        return nn.ModuleList([actnorm()])
    layers = []
    for _ in range(num_flows):
        layers += [actnorm(), InvertibleConv(c_in=num_dims),
                   CouplingLayer(c_in=num_dims, mask=CouplingLayer.create_channel_mask(c_in=num_dims),
                                 block_type="LinearNet", model_func=coupling_net)]
    return nn.ModuleList(layers)
