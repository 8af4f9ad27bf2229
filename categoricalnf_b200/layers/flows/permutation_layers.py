"""Invertible 1x1 convolution (reference layers/flows/permutation_layers.py:11-140).

Same parameters / buffers as upstream (``p``, ``sign_s``, ``l``, ``log_s``, ``u``, ``l_mask``,
``eye`` or ``weight``).  W, its float64 inverse and log|det| are produced by one single-CTA kernel
(``cnf_invconv_build``); the per-position product by ``cnf_invconv_apply`` (csrc/invconv.cu).
In eval mode the built matrices are cached in ``eval_dict`` like upstream, but keyed on the
parameter version counters and ``ops.param_epoch()`` so a ``load_state_dict`` or an optimiser step
(also one that writes through ``p.data``) invalidates them (fixes App. B #12).
"""
from collections import defaultdict

import numpy as np
import scipy.linalg
import torch
import torch.nn as nn

from ... import functional as CF
from ... import ops
from .flow_layer import FlowLayer


def _initial_weight(c_in):
    """Random rotation: for two channels an angle away from the identity / a flip, otherwise the Q
    factor of a Gaussian matrix (:17-34).  Draws from ``np.random`` like upstream."""
    if c_in == 2:
        r = np.random.uniform()
        angle = (0.25 + r) * np.pi if r < 0.5 else (1.25 + (2 * r - 1) * 0.5) * np.pi
        return np.array([[np.cos(angle), -np.sin(angle)], [np.sin(angle), np.cos(angle)]])
    return np.linalg.qr(np.random.randn(c_in, c_in))[0].astype(np.float32)


class InvertibleConv(FlowLayer):

    def __init__(self, c_in, LU_decomposed=True):
        super().__init__()
        self.num_channels = c_in
        self.LU_decomposed = LU_decomposed
        w0 = _initial_weight(c_in)
        if not LU_decomposed:
            self.weight = nn.Parameter(torch.from_numpy(np.asarray(w0, dtype=np.float32)))
        else:
            perm, lower, upper = scipy.linalg.lu(w0)
            diag = np.diag(upper)
            f32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
            self.register_buffer("p", f32(perm))
            self.register_buffer("sign_s", f32(np.sign(diag)))
            self.l = nn.Parameter(f32(lower))
            self.log_s = nn.Parameter(f32(np.log(np.abs(diag))))
            self.u = nn.Parameter(f32(np.triu(upper, k=1)))
            self.register_buffer("l_mask", torch.tril(torch.ones(c_in, c_in), -1))
            self.register_buffer("eye", torch.eye(c_in))
        self.eval_dict = defaultdict(self._get_default_inner_dict)

    def _get_default_inner_dict(self):
        return {"weight": None, "inv_weight": None, "sldj": None, "version": None}

    def _param_version(self):
        ps = [self.weight] if not self.LU_decomposed else [self.p, self.sign_s, self.l, self.log_s, self.u]
        return (ops.param_epoch(),) + tuple((t.data_ptr(), t._version) for t in ps)

    def _build(self, differentiable):
        if differentiable:
            # training: W and sum(log_s) must carry gradients to l / u / log_s - tiny C x C eager ops
            if not self.LU_decomposed:
                w = self.weight
                return w, None, torch.slogdet(w)[1]
            lo = self.l * self.l_mask + self.eye
            up = self.u * self.l_mask.t() + torch.diag(self.sign_s * torch.exp(self.log_s))
            return self.p @ (lo @ up), None, self.log_s.sum()
        if not self.LU_decomposed:
            return ops.invconv_build(weight=self.weight.detach())
        return ops.invconv_build(p=self.p, l=self.l.detach(), u=self.u.detach(), log_s=self.log_s.detach(),
                                 sign_s=self.sign_s)

    def _get_weight(self, device_name, inverse=False):
        """(W or W^-1, sum log|s|) - W^-1 is the float64 inverse rounded to float32 (:77, :85)."""
        needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        if self.training or needs_grad:
            self.eval_dict.pop(device_name, None)
            w, w_inv, sldj = self._build(differentiable=needs_grad)
            if inverse and w_inv is None:
                w_inv = torch.inverse(w.double()).float() if needs_grad else ops.invconv_build(weight=w.detach())[1]
            return (w_inv if inverse else w), sldj
        entry = self.eval_dict[device_name]
        if entry["weight"] is None or entry["version"] != self._param_version():
            w, w_inv, sldj = self._build(differentiable=False)
            entry.update(weight=w, inv_weight=w_inv, sldj=sldj, version=self._param_version())
        return (entry["inv_weight"] if inverse else entry["weight"]), entry["sldj"]

    def _is_eval_dict_empty(self, device_name=None):
        return (device_name not in self.eval_dict) if device_name is not None else len(self.eval_dict) == 0

    def _empty_eval_dict(self, device_name=None):
        if device_name is not None:
            self.eval_dict.pop(device_name, None)
        else:
            self.eval_dict = defaultdict(self._get_default_inner_dict)

    def forward(self, x, ldj=None, reverse=False, length=None, channel_padding_mask=None, layer_share_dict=None,
                **kwargs):
        if ldj is None:
            ldj = x.new_zeros(x.size(0))
        else:
            ldj = ldj.clone()       # upstream builds a new tensor here (`ldj = ldj + sldj`, :114-117)
        weight, sldj = self._get_weight(device_name=str(x.device), inverse=reverse)
        length = None if length is None else length.float()
        z, ldj = CF.invconv(x, weight, sldj, ldj, pad=channel_padding_mask, length=length, reverse=reverse)
        if layer_share_dict is not None:
            for key in ("t", "log_s", "error_decay"):
                if key in layer_share_dict:
                    layer_share_dict[key] = layer_share_dict[key] * 0.0
        return z, ldj

    def info(self):
        return "Invertible 1x1 Convolution - %i channels %s" % (self.num_channels,
                                                               "(LU decomposed)" if self.LU_decomposed else "")
