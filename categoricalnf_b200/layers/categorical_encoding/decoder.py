"""Embedding factory and the linear decoder of the categorical encodings
(reference layers/categorical_encoding/decoder.py:11-63).  The decoder MLP is a ``LinearNet`` of ``TCLinear`` layers (the
tcgen05 projection kernel, GELU in its epilogue) with the reference's parameter names (``layers.inp_layer.0.*``,
``layers.main_net.<i>.*``).
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from ..networks.help_layers import LinearNet


def _param(config, key, default):
    val = config.get(key, default) if config is not None else default
    return default if val is None else val


def create_embed_layer(vocab, vocab_size, default_embed_layer_dims):
    """``nn.Embedding`` from a torchtext-style vocabulary (pre-trained vectors) or from scratch (:11-20)."""
    pretrained = vocab is not None and getattr(vocab, "vectors", None) is not None
    dims = vocab.vectors.shape[1] if pretrained else default_embed_layer_dims
    vocab_size = len(vocab) if pretrained else vocab_size
    layer = nn.Embedding(vocab_size, dims)
    if pretrained:
        layer.weight.data.copy_(vocab.vectors)
        layer.weight.requires_grad = True
    return layer, vocab_size


def create_decoder(num_categories, num_dims, config, **kwargs):
    return DecoderLinear(num_categories, embed_dim=num_dims, hidden_size=_param(config, "hidden_size", 64),
                         num_layers=_param(config, "num_layers", 1), **kwargs)


class DecoderLinear(nn.Module):
    """MLP over ``[z, elu(z), elu(-z)]`` with a log-softmax over the categories (:35-63)."""

    def __init__(self, num_categories, embed_dim, hidden_size, num_layers, class_prior_log=None):
        super().__init__()
        self.hidden_size, self.num_layers = hidden_size, num_layers
        self.layers = LinearNet(c_in=3 * embed_dim, c_out=num_categories, hidden_size=hidden_size, num_layers=num_layers)
        self.log_softmax = nn.LogSoftmax(dim=-1)
        if class_prior_log is not None:
            if isinstance(class_prior_log, np.ndarray):
                class_prior_log = torch.from_numpy(class_prior_log)
            self.layers.set_bias(class_prior_log)

    def forward(self, z_cont):
        feats = torch.cat([z_cont, F.elu(z_cont), F.elu(-z_cont)], dim=-1)
        return self.log_softmax(self.layers(feats))

    def info(self):
        return "Linear model with hidden size %i and %i layers" % (self.hidden_size, self.num_layers)
