"""Autograd-aware functional layer between the drop-in modules and :mod:`categoricalnf_b200.ops`.

Each function is one fused kernel launch in the forward direction.  Functions that sit on a
training path are wrapped in ``torch.autograd.Function`` so that the reference's unchanged
training loops (``loss.backward()``, general/train.py:148-152) differentiate through them with the
hand-written backward kernels.
"""
from __future__ import annotations

import os

import torch

from . import ops


def _needs_grad(*tensors) -> bool:
    return torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors)


# ----------------------------------------------------------------------------------------------
# mixture-CDF coupling
# ----------------------------------------------------------------------------------------------
class _MixCDF(torch.autograd.Function):

    @staticmethod
    def forward(ctx, z, nn_out, sf, msf, pad, cfg):
        z_out, ldj, reg = ops.mixcdf(z, nn_out, cfg["K"], mask_c=cfg["mask_c"], mask_s=cfg["mask_s"], pad=pad,
                                     scaling_factor=sf, mixture_scaling_factor=msf, reverse=cfg["reverse"],
                                     reg_max=cfg["reg_max"], reg_factor=cfg["reg_factor"], training=cfg["training"],
                                     want_reg=True, prebounded=cfg.get("prebounded", False), compact=cfg.get("compact", False))
        ctx.cfg = cfg
        ctx.save_for_backward(z, nn_out, sf, msf, pad, z_out)
        ctx.mark_non_differentiable(reg)
        return z_out, ldj, reg

    @staticmethod
    def backward(ctx, g_z, g_ldj, g_reg):
        from . import ops_bwd
        z, nn_out, sf, msf, pad, z_out = ctx.saved_tensors
        gz, gnn, gsf, gmsf = ops_bwd.mixcdf_backward(ctx.cfg, z, nn_out, sf, msf, pad, z_out, g_z, g_ldj,
                                                     ctx.needs_input_grad)
        return gz, gnn, gsf, gmsf, None, None


def mixcdf(z, nn_out, num_mixtures, scaling_factor=None, mixture_scaling_factor=None, *, mask_c=None, mask_s=None,
           pad=None, reverse=False, reg_max=-1.0, reg_factor=1.0, training=False, prebounded=False, compact=False):
    """(z_out, ldj [B], reg_ldj [B]) of the logistic-mixture coupling transform (K1 / K2).  ``compact``: ``nn_out`` (and
    its gradient) hold the transformed channels' records only, [B,S,Ct*(2+3K)]."""
    cfg = dict(K=int(num_mixtures), mask_c=mask_c, mask_s=mask_s, reverse=bool(reverse), reg_max=float(reg_max),
               reg_factor=float(reg_factor), training=bool(training), prebounded=bool(prebounded), compact=bool(compact))
    if _needs_grad(z, nn_out, scaling_factor, mixture_scaling_factor):
        return _MixCDF.apply(z, nn_out, scaling_factor, mixture_scaling_factor, pad, cfg)
    return ops.mixcdf(z, nn_out, cfg["K"], mask_c=mask_c, mask_s=mask_s, pad=pad, scaling_factor=scaling_factor,
                      mixture_scaling_factor=mixture_scaling_factor, reverse=reverse, reg_max=reg_max,
                      reg_factor=reg_factor, training=training, want_reg=True, prebounded=prebounded, compact=compact)


class _ProjMixCDF(torch.autograd.Function):
    """Final projection of the coupling network (compact form: only the transformed channels' weight rows) + mixture
    transform as ONE autograd node, so that the backward kernel of the transform can hand the projection's bias gradient
    over: it is the column sum of dL/dnn_out, which ``cnf_mixcdf_bwd`` (ABI v5 ``grad_nn_colsum``) forms while the
    gradient tile is still in shared memory - the separate column-sum pass re-read the whole gradient (0.87 GB at the LM
    shape).  Forward and the other gradients are exactly ``_TCLinearFn`` followed by ``_MixCDF``."""

    @staticmethod
    def forward(ctx, z, feats, weight, bias, sf, msf, pad, cfg, precision):
        B, S = z.shape[0], z.shape[1]
        split = None
        if precision == "3xtf32" and weight.dtype == torch.float32 and weight.is_contiguous():
            split = ops.weight_split(weight, use_cache=False)
        ctx.has_split = split is not None
        nn_out = ops.linear(feats, weight, bias, precision=precision, cache_weight=False, split=split).view(B, S, weight.shape[0])
        z_out, ldj, reg = ops.mixcdf(z, nn_out, cfg["K"], mask_c=cfg["mask_c"], mask_s=cfg["mask_s"], pad=pad,
                                     scaling_factor=sf, mixture_scaling_factor=msf, reverse=False,
                                     reg_max=cfg["reg_max"], reg_factor=cfg["reg_factor"], training=cfg["training"],
                                     want_reg=True, compact=True)
        ctx.cfg, ctx.precision, ctx.has_bias = cfg, precision, bias is not None
        # the network is one per-position Linear on z itself: its backward products can run inside the transform's backward
        # kernel (cnf_mixcdf_bwd proj_weight, compiled for the LM layout: 16 channels, 8 contiguous transformed ones)
        mc = cfg["mask_c"]
        tch = [c for c, m in enumerate(mc)] if mc is None else [c for c, m in enumerate(mc) if float(m) == 0.0]
        ctx.fuse_linear_bwd = bool(
            cfg.get("linear_on_z", False) and FUSE_LINEAR_BACKWARD and feats.data_ptr() == z.data_ptr() and z.shape[-1] == 16
            and feats.shape[-1] == 16 and len(tch) == 8 and tch in (list(range(0, 8)), list(range(8, 16))) and cfg["K"] in (4, 8, 16)
            and weight.is_contiguous())
        ctx.save_for_backward(z, nn_out, sf, msf, pad, z_out, feats, weight, *(split or ()))
        ctx.mark_non_differentiable(reg)
        return z_out, ldj, reg

    @staticmethod
    def backward(ctx, g_z, g_ldj, g_reg):
        from . import ops_bwd
        from .layers.networks.linear import BACKWARD_PRECISION
        z, nn_out, sf, msf, pad, z_out, feats, weight = ctx.saved_tensors[:8]
        split = tuple(ctx.saved_tensors[8:10]) if ctx.has_split else None
        need = ctx.needs_input_grad
        want_col = ctx.has_bias and need[3]
        if ctx.fuse_linear_bwd:
            gz, _, gsf, gmsf, gcol, gw = ops_bwd.mixcdf_backward(ctx.cfg, z, nn_out, sf, msf, pad, z_out, g_z, g_ldj, need,
                                                                  want_colsum=want_col, proj_weight=weight,
                                                                  want_proj_weight_grad=need[2])
            # (the gradient with respect to `feats` - the same tensor as z - is already inside gz)
            return gz, None, gw, gcol, gsf, gmsf, None, None, None
        out = ops_bwd.mixcdf_backward(ctx.cfg, z, nn_out, sf, msf, pad, z_out, g_z, g_ldj, need, want_colsum=want_col)
        gz, gnn, gsf, gmsf = out[:4]
        gx, gw, _ = ops.linear_bwd(feats, weight, gnn.view(-1, gnn.shape[-1]), need_x=need[1], need_weight=need[2], need_bias=False,
                                   precision=BACKWARD_PRECISION or ctx.precision, weight_split=split)
        return gz, gx, gw, (out[4] if want_col else None), gsf, gmsf, None, None, None


FUSE_LINEAR_BACKWARD = os.environ.get("CNF_B200_NO_FUSED_LINEAR_BWD", "0") in ("", "0")      # A/B switch


def proj_mixcdf(z, feats, weight, bias, num_mixtures, scaling_factor=None, mixture_scaling_factor=None, *, mask_c=None,
                pad=None, reg_max=-1.0, reg_factor=1.0, training=False, precision="3xtf32", linear_on_z=False):
    """``mixcdf(z, feats @ weight.T + bias, ..., compact=True)`` (forward direction) with ``weight`` / ``bias`` holding the
    transformed channels' rows only; ``feats`` [B*S, H].  -> (z_out, ldj [B], reg_ldj [B]).  ``linear_on_z``: ``feats`` is z
    itself and ``weight`` already carries the coupling mask (zero columns for the transformed inputs) - the Linear's backward
    then runs inside the transform's backward kernel where that is compiled."""
    cfg = dict(K=int(num_mixtures), mask_c=mask_c, mask_s=None, reverse=False, reg_max=float(reg_max),
               reg_factor=float(reg_factor), training=bool(training), prebounded=False, compact=True, linear_on_z=bool(linear_on_z))
    return _ProjMixCDF.apply(z, feats, weight, bias, scaling_factor, mixture_scaling_factor, pad, cfg, precision)


# ----------------------------------------------------------------------------------------------
# affine coupling
# ----------------------------------------------------------------------------------------------
class _Affine(torch.autograd.Function):

    @staticmethod
    def forward(ctx, z, nn_out, sf, cfg):
        ldj = torch.zeros(z.size(0), dtype=torch.float32, device=z.device)
        z_out, ldj = ops.affine_coupling(z, nn_out, ldj, mask_c=cfg["mask_c"], mask_s=cfg["mask_s"],
                                         scaling_factor=sf, reverse=cfg["reverse"], prebounded=cfg["prebounded"])
        ctx.cfg = cfg
        ctx.save_for_backward(z, nn_out, sf, z_out)
        return z_out, ldj

    @staticmethod
    def backward(ctx, g_z, g_ldj):
        from . import ops_bwd
        z, nn_out, sf, z_out = ctx.saved_tensors
        gz, gnn, gsf = ops_bwd.affine_backward(ctx.cfg, z, nn_out, sf, z_out, g_z, g_ldj, ctx.needs_input_grad)
        return gz, gnn, gsf, None


def affine_coupling(z, nn_out, scaling_factor=None, *, mask_c=None, mask_s=None, reverse=False, prebounded=False):
    """(z_out, layer_ldj [B]) of the affine coupling transform (K3)."""
    cfg = dict(mask_c=mask_c, mask_s=mask_s, reverse=bool(reverse), prebounded=bool(prebounded))
    if _needs_grad(z, nn_out, scaling_factor):
        return _Affine.apply(z, nn_out, scaling_factor, cfg)
    ldj = torch.zeros(z.size(0), dtype=torch.float32, device=z.device)
    return ops.affine_coupling(z, nn_out, ldj, mask_c=mask_c, mask_s=mask_s, scaling_factor=scaling_factor,
                               reverse=reverse, prebounded=prebounded)


def affine_explicit(z, s, t, reverse=False):
    """``CouplingLayer.run_with_params`` for explicit bounded (s, t): packs them into the [s, t]
    record layout (pure data movement) and runs the same kernel with bounding disabled."""
    rec = torch.stack([s.expand_as(z), t.expand_as(z)], dim=-1).reshape(z.shape[:-1] + (2 * z.shape[-1],))
    return affine_coupling(z, rec, None, reverse=reverse, prebounded=True)


# ----------------------------------------------------------------------------------------------
# activation normalisation / 1x1 convolution / prior: ldj is updated IN PLACE like upstream
# ----------------------------------------------------------------------------------------------
class _ActNorm(torch.autograd.Function):

    @staticmethod
    def forward(ctx, z, bias, scales, ldj, pad, length, reverse):
        z_out, _ = ops.actnorm(z, bias, scales, ldj, pad=pad, length=length, reverse=reverse)
        ctx.mark_dirty(ldj)
        ctx.reverse = reverse
        ctx.save_for_backward(z, bias, scales, pad, length, z_out)
        return z_out, ldj

    @staticmethod
    def backward(ctx, g_z, g_ldj):
        from . import ops_bwd
        z, bias, scales, pad, length, z_out = ctx.saved_tensors
        gz, gb, gs = ops_bwd.actnorm_backward(z, bias, scales, pad, length, z_out, g_z, g_ldj, ctx.reverse,
                                              ctx.needs_input_grad)
        return gz, gb, gs, g_ldj, None, None, None


def actnorm(z, bias, scales, ldj, *, pad=None, length=None, reverse=False):
    if _needs_grad(z, bias, scales, ldj):
        return _ActNorm.apply(z, bias, scales, ldj, pad, length, bool(reverse))
    return ops.actnorm(z, bias, scales, ldj, pad=pad, length=length, reverse=reverse)


class _ExtActNorm(torch.autograd.Function):

    @staticmethod
    def forward(ctx, z, ext, ldj, pad, reverse):
        z_out, _ = ops.ext_actnorm(z, ext, ldj, pad=pad, reverse=reverse)
        ctx.mark_dirty(ldj)
        ctx.reverse = reverse
        ctx.save_for_backward(z, ext, pad, z_out)
        return z_out, ldj

    @staticmethod
    def backward(ctx, g_z, g_ldj):
        from . import ops_bwd
        z, ext, pad, z_out = ctx.saved_tensors
        gz, gext = ops_bwd.ext_actnorm_backward(z, ext, pad, z_out, g_z, g_ldj, ctx.reverse, ctx.needs_input_grad)
        return gz, gext, g_ldj, None, None


def ext_actnorm(z, ext, ldj, *, pad=None, reverse=False):
    if _needs_grad(z, ext, ldj):
        return _ExtActNorm.apply(z, ext, ldj, pad, bool(reverse))
    return ops.ext_actnorm(z, ext, ldj, pad=pad, reverse=reverse)


class _InvConv(torch.autograd.Function):

    @staticmethod
    def forward(ctx, z, weight, sldj, ldj, pad, length, reverse):
        z_out, _ = ops.invconv_apply(z, weight, sldj, ldj, pad=pad, length=length, reverse=reverse)
        ctx.mark_dirty(ldj)
        ctx.reverse = reverse
        ctx.sldj_shape = sldj.shape
        ctx.save_for_backward(z, weight, pad, length)
        return z_out, ldj

    @staticmethod
    def backward(ctx, g_z, g_ldj):
        from . import ops_bwd
        z, weight, pad, length = ctx.saved_tensors
        gz, gw, gsldj = ops_bwd.invconv_backward(z, weight, pad, length, g_z, g_ldj, ctx.reverse, ctx.needs_input_grad)
        return gz, gw, gsldj.reshape(ctx.sldj_shape), g_ldj, None, None, None


def invconv(z, weight, sldj, ldj, *, pad=None, length=None, reverse=False):
    if _needs_grad(z, weight, sldj, ldj):
        return _InvConv.apply(z, weight, sldj, ldj, pad, length, bool(reverse))
    return ops.invconv_apply(z, weight, sldj, ldj, pad=pad, length=length, reverse=reverse)


class _LogisticLogProb(torch.autograd.Function):

    @staticmethod
    def forward(ctx, x, mu, sigma):
        _, lp = ops.logistic_logprob(x, mu=mu, sigma=sigma, reduce=False, elementwise=True)
        ctx.mu, ctx.sigma = mu, sigma
        ctx.save_for_backward(x)
        return lp

    @staticmethod
    def backward(ctx, g):
        from . import ops_bwd
        (x,) = ctx.saved_tensors
        return ops_bwd.logistic_logprob_backward(x, g, ctx.mu, ctx.sigma), None, None


def logistic_logprob(x, mu=0.0, sigma=1.0 / 1.81):
    """Element-wise log-density of Logistic(mu, sigma) (K7)."""
    if _needs_grad(x):
        return _LogisticLogProb.apply(x, float(mu), float(sigma))
    return ops.logistic_logprob(x, mu=mu, sigma=sigma, reduce=False, elementwise=True)[1]


# ----------------------------------------------------------------------------------------------
# SigmoidFlow (sigmoid_layer.py:24-48)
# ----------------------------------------------------------------------------------------------
class _SigmoidFlow(torch.autograd.Function):

    @staticmethod
    def forward(ctx, z, ldj, reverse, alpha, sum_ldj, add_tokens):
        z_out, l = ops.sigmoid_flow(z, None, reverse=reverse, alpha=alpha, sum_ldj=sum_ldj, add_tokens=add_tokens)
        ctx.reverse, ctx.alpha, ctx.sum_ldj = reverse, alpha, sum_ldj
        ctx.save_for_backward(z)
        if sum_ldj and ldj is not None:
            l = l + ldj
        return z_out, l

    @staticmethod
    def backward(ctx, g_z, g_l):
        from . import ops_bwd
        (z,) = ctx.saved_tensors
        gz = ops_bwd.sigmoid_flow_backward(z, g_z, g_l if ctx.sum_ldj else None, None if ctx.sum_ldj else g_l,
                                           ctx.reverse, ctx.alpha)
        return gz, (g_l if ctx.sum_ldj else None), None, None, None, None


def sigmoid_flow(z, ldj=None, *, reverse=False, alpha=1e-5, sum_ldj=True, add_tokens=None):
    """(z_out, ldj) of SigmoidFlow in its effective direction; ``ldj`` is a new tensor (``ldj + layer_ldj``, :44)."""
    if _needs_grad(z, ldj):
        return _SigmoidFlow.apply(z, ldj, bool(reverse), float(alpha), bool(sum_ldj), add_tokens)
    return ops.sigmoid_flow(z, ldj, reverse=reverse, alpha=alpha, sum_ldj=sum_ldj, add_tokens=add_tokens)


# ----------------------------------------------------------------------------------------------
# mixture-of-logistics categorical encoding: differentiable in the class table
# ----------------------------------------------------------------------------------------------
class _CategEncode(torch.autograd.Function):

    @staticmethod
    def forward(ctx, table, tokens, category_prior, pad, cfg):
        ldj = torch.zeros(tokens.shape[0], dtype=torch.float32, device=tokens.device)
        z, ldj, cpl = ops.categ_encode(tokens, table, category_prior, ldj, noise=cfg["noise"], seed=cfg["seed"], offset=cfg["offset"],
                                       pad=pad, beta=cfg["beta"], want_class_prob=True)
        ctx.beta = cfg["beta"]
        ctx.save_for_backward(tokens, z, table, category_prior, pad)
        ctx.mark_non_differentiable(cpl)
        return z, ldj, cpl

    @staticmethod
    def backward(ctx, g_z, g_ldj, g_cpl):
        from . import ops_bwd
        tokens, z, table, prior, pad = ctx.saved_tensors
        gtable = ops_bwd.categ_encode_backward(tokens, z, table, prior, pad, ctx.beta, g_z, g_ldj)
        return gtable, None, None, None, None


def categ_encode(tokens, table, category_prior, *, noise=None, seed=0, offset=0, pad=None, beta=1.0):
    """(z [B,S,D], ldj [B], class_prob_log [B,S]) of the mixture-of-logistics encoding (K6); differentiable with respect to
    ``table`` [V,2D] (the noise is a constant of the graph, as in the reference where it is a fresh sample)."""
    cfg = dict(noise=noise, seed=int(seed), offset=int(offset), beta=float(beta))
    if _needs_grad(table):
        return _CategEncode.apply(table, tokens, category_prior, pad, cfg)
    ldj = torch.zeros(tokens.shape[0], dtype=torch.float32, device=tokens.device)
    return ops.categ_encode(tokens, table, category_prior, ldj, noise=noise, seed=seed, offset=offset, pad=pad, beta=beta,
                            want_class_prob=True)
