"""Activation normalisation layers (reference layers/flows/activation_normalization.py).

``ActNormFlow``    learned per-channel bias / log-scale with data-dependent init
``ExtActNormFlow`` bias / log-scale predicted from an external input (the class embedding) - the
                   class-conditional logistic of the mixture encoding

Both update the caller's ``ldj`` tensor in place, exactly like upstream's ``ldj += ...``
(App. B #2).  One kernel launch each (csrc/elementwise.cu).
"""
import torch
import torch.nn as nn

from ... import functional as CF
from ... import ops
from .flow_layer import FlowLayer


class ActNormFlow(FlowLayer):

    def __init__(self, c_in, data_init=True):
        super().__init__()
        self.c_in = c_in
        self.data_init = data_init
        self.bias = nn.Parameter(torch.zeros(1, 1, self.c_in))
        self.scales = nn.Parameter(torch.zeros(1, 1, self.c_in))

    def forward(self, z, ldj=None, reverse=False, length=None, channel_padding_mask=None, **kwargs):
        if ldj is None:
            ldj = z.new_zeros(z.size(0))
        # per-sample length: explicit `length`, else the number of unpadded positions, else S
        # (:27-33) - the last two are resolved inside the kernel
        length = None if length is None else length.float()
        z, ldj = CF.actnorm(z, self.bias, self.scales, ldj, pad=channel_padding_mask, length=length, reverse=reverse)
        return z, ldj

    def need_data_init(self):
        return self.data_init

    def data_init_forward(self, input_data, channel_padding_mask=None, **kwargs):
        """Set bias / scales so the (unpadded) activations have zero mean and unit variance (:55-67)."""
        with torch.no_grad():
            bias, scales = ops.actnorm_data_init(input_data, channel_padding_mask)
            self.bias.data = bias.view(1, 1, -1)
            self.scales.data = scales.view(1, 1, -1)
            out, _ = ops.actnorm(input_data, self.bias, self.scales, None, pad=channel_padding_mask)
            n = out.shape[0] * out.shape[1] if channel_padding_mask is None else channel_padding_mask.sum()
            mean = out.sum(dim=[0, 1]) / n
            print("[INFO - ActNorm] New mean", mean)
            print("[INFO - ActNorm] New variance", torch.sqrt((out ** 2).sum(dim=[0, 1]) / n - mean ** 2))

    def info(self):
        return "Activation Normalizing Flow (c_in=%i)" % self.c_in


class ExtActNormFlow(FlowLayer):

    def __init__(self, c_in, net, zero_init=False, data_init=False, make_unique=False):
        super().__init__()
        self.c_in = c_in
        self.data_init = data_init
        self.make_unique = make_unique
        self.pred_net = net
        if zero_init:
            if hasattr(self.pred_net, "initialize_zeros"):
                self.pred_net.initialize_zeros()
            elif isinstance(self.pred_net, nn.Sequential):
                self.pred_net[-1].weight.data.zero_()
                self.pred_net[-1].bias.data.zero_()

    def _run_nn(self, ext_input):
        if not self.make_unique:
            return self.pred_net(ext_input)
        # evaluate the predictor once per distinct input value and scatter back (:103-113)
        uniq, inverse = torch.unique(ext_input, return_inverse=True)
        outs = self.pred_net(uniq)
        return outs.index_select(0, inverse.reshape(-1)).reshape(ext_input.shape + outs.shape[-1:])

    def forward(self, z, ldj=None, reverse=False, ext_input=None, channel_padding_mask=None, layer_share_dict=None,
                **kwargs):
        if ldj is None:
            ldj = z.new_zeros(z.size(0))
        if ext_input is None:
            print("[!] WARNING: External input in ExtActNormFlow is None. Using default params...")
            nn_out = z.new_zeros(z.shape[:-1] + (2 * z.shape[-1],))
        else:
            nn_out = self._run_nn(ext_input)
        z_out, ldj = CF.ext_actnorm(z, nn_out, ldj, pad=channel_padding_mask, reverse=reverse)
        if layer_share_dict is not None and not reverse:
            bias, scales = nn_out.chunk(2, dim=2)
            scales = torch.tanh(scales)
            layer_share_dict["t"] = (layer_share_dict["t"] + bias) * torch.exp(scales)
            layer_share_dict["log_s"] = layer_share_dict["log_s"] + scales
        return z_out, ldj

    def need_data_init(self):
        return self.data_init

    def data_init_forward(self, input_data, channel_padding_mask=None, **kwargs):
        """Data-dependent init of the predictor's output bias (:150-171).  Upstream sums *all*
        elements (padded included) but divides by the number of unpadded ones; kept as is."""
        with torch.no_grad():
            mask = input_data.new_ones(input_data.shape) if channel_padding_mask is None else \
                channel_padding_mask.view(input_data.shape[:-1] + channel_padding_mask.shape[-1:])
            n = mask.sum(dim=[0, 1], keepdim=True)
            bias = -input_data.sum(dim=[0, 1], keepdim=True) / n
            scale = -0.5 * ((((input_data + bias) ** 2) * mask).sum(dim=[0, 1], keepdim=True) / n).log()
            packed = torch.cat([bias, scale], dim=-1).squeeze()
            if isinstance(self.pred_net, nn.Sequential):
                self.pred_net[-1].bias.data = packed
            else:
                self.pred_net.set_bias(packed)

    def info(self):
        return "External Activation Normalizing Flow (c_in=%i)" % self.c_in
