"""Backward passes of the fused layers: thin tensor-level wrappers over the
``cnf_*_bwd`` kernels, called by the ``torch.autograd.Function``s in :mod:`categoricalnf_b200.functional`.
Like the forward ops there is no eager fallback: CPU tensors are rejected.
"""
from __future__ import annotations

import torch

from . import _lib as L
from . import ops
from .ops import _call, _f32, _mask_struct, _opt_f32, _pad_bs, _ptr


def _grad(t, like):
    """dense float32 gradient, zeros when autograd passed None"""
    if t is None:
        return torch.zeros_like(like, dtype=torch.float32)
    return _f32(t, "grad")


def mixcdf_backward(cfg, z, nn_out, sf, msf, pad, z_out, g_z, g_ldj, needs, want_colsum=False, proj_weight=None,
                    want_proj_weight_grad=False):
    """-> (grad_z, grad_nn_out, grad_sf | None, grad_msf | None[, column sums of grad_nn_out]).  ``want_colsum`` (compact layout
    only, ABI v5): the bias gradient of the network's final Linear, summed by the backward kernel itself.  ``proj_weight``
    [Ct*(2+3K), C] (ABI v5, see ``cnf_mixcdf_bwd_args.proj_weight``): the network is one per-position Linear on z - the kernel
    adds its input gradient into grad_z, returns its weight gradient (``want_proj_weight_grad``) as a sixth value and does
    NOT write grad_nn_out (returned as None)."""
    if cfg["reverse"]:
        raise NotImplementedError("categoricalnf_b200: differentiating the INVERSE mixture coupling is not supported "
                                  "(training differentiates the density direction only)")
    z = _f32(z, "z")
    B, S, Cc = z.shape
    K = cfg["K"]
    compact = bool(cfg.get("compact", False))
    nn_out = _f32(nn_out, "nn_out", (B, S, (ops._n_transformed(cfg["mask_c"], Cc) if compact else Cc) * (2 + 3 * K)))
    a = L.MixcdfBwdArgs()
    a.B, a.S, a.C, a.K = B, S, Cc, K
    a.mask, keep = _mask_struct(cfg["mask_c"], cfg["mask_s"])
    pad = _pad_bs(pad, B, S)
    sf = _opt_f32(sf, "scaling_factor", (Cc,))
    msf = _opt_f32(msf, "mixture_scaling_factor", (Cc, K))
    gz_out = _grad(g_z, z)
    gl = _opt_f32(g_ldj, "grad_ldj", (B,))
    gz = torch.empty_like(z)
    gnn = torch.empty_like(nn_out) if proj_weight is None else None
    pre = bool(cfg.get("prebounded", False))
    gsf = torch.zeros(Cc, dtype=torch.float32, device=z.device) if (sf is not None and not pre) else None
    gmsf = torch.zeros(Cc, K, dtype=torch.float32, device=z.device) if (msf is not None and not pre) else None
    a.z, a.nn_out, a.pad, a.scaling_factor, a.mixture_scaling_factor = _ptr(z), _ptr(nn_out), _ptr(pad), _ptr(sf), _ptr(msf)
    a.reg_max, a.reg_factor, a.training = float(cfg["reg_max"]), float(cfg["reg_factor"]), int(bool(cfg["training"]))
    a.params_prebounded = int(pre)
    a.nn_compact = int(compact)
    a.grad_z_out, a.grad_ldj, a.grad_z, a.grad_nn_out = _ptr(gz_out), _ptr(gl), _ptr(gz), _ptr(gnn)
    a.grad_scaling_factor, a.grad_mixture_scaling_factor = _ptr(gsf), _ptr(gmsf)
    gcol = None
    if want_colsum:
        if not compact:
            raise ValueError("mixcdf_backward: column sums need the compact layout")
        gcol = torch.zeros(nn_out.shape[-1], dtype=torch.float32, device=z.device)
        a.grad_nn_colsum = _ptr(gcol)
    gpw = None
    if proj_weight is not None:
        if not compact:
            raise ValueError("mixcdf_backward: proj_weight needs the compact layout")
        proj_weight = _f32(proj_weight, "proj_weight", (nn_out.shape[-1], Cc))
        a.proj_weight = _ptr(proj_weight)
        if want_proj_weight_grad:
            gpw = torch.zeros_like(proj_weight)
            a.grad_proj_weight = _ptr(gpw)
    _call("cnf_mixcdf_bwd", a, z, (keep, z, nn_out, pad, sf, msf, gz_out, gl, proj_weight))
    if proj_weight is not None:
        return gz, gnn, gsf, gmsf, gcol, gpw
    if want_colsum:
        return gz, gnn, gsf, gmsf, gcol
    return gz, gnn, gsf, gmsf


def affine_backward(cfg, z, nn_out, sf, z_out, g_z, g_ldj, needs):
    z = _f32(z, "z")
    B, S, Cc = z.shape
    nn_out = _f32(nn_out, "nn_out", (B, S, 2 * Cc))
    a = L.AffineBwdArgs()
    a.B, a.S, a.C = B, S, Cc
    a.mask, keep = _mask_struct(cfg["mask_c"], cfg["mask_s"])
    sf = _opt_f32(sf, "scaling_factor", (Cc,))
    gz_out = _grad(g_z, z)
    gl = _opt_f32(g_ldj, "grad_ldj", (B,))
    gz, gnn = torch.empty_like(z), torch.empty_like(nn_out)
    pre = bool(cfg["prebounded"])
    gsf = torch.zeros(Cc, dtype=torch.float32, device=z.device) if (sf is not None and not pre) else None
    a.z, a.nn_out, a.scaling_factor, a.reverse, a.params_prebounded = _ptr(z), _ptr(nn_out), _ptr(sf), int(cfg["reverse"]), int(pre)
    a.grad_z_out, a.grad_ldj, a.grad_z, a.grad_nn_out, a.grad_scaling_factor = _ptr(gz_out), _ptr(gl), _ptr(gz), _ptr(gnn), _ptr(gsf)
    _call("cnf_affine_coupling_bwd", a, z, (keep, z, nn_out, sf, gz_out, gl))
    return gz, gnn, gsf


def actnorm_backward(z, bias, scales, pad, length, z_out, g_z, g_ldj, reverse, needs):
    z = _f32(z, "z")
    B, S, Cc = z.shape
    b1, s1 = _f32(bias, "bias").reshape(-1), _f32(scales, "scales").reshape(-1)
    pad = _pad_bs(pad, B, S)
    length = _opt_f32(length, "length", (B,))
    gz_out = _grad(g_z, z)
    gl = _opt_f32(g_ldj, "grad_ldj", (B,))
    gz = torch.empty_like(z)
    gb = torch.zeros(Cc, dtype=torch.float32, device=z.device)
    gs = torch.zeros(Cc, dtype=torch.float32, device=z.device)
    a = L.ActnormBwdArgs()
    a.B, a.S, a.C = B, S, Cc
    a.z, a.bias, a.scales, a.pad, a.length, a.reverse = _ptr(z), _ptr(b1), _ptr(s1), _ptr(pad), _ptr(length), int(bool(reverse))
    a.grad_z_out, a.grad_ldj, a.grad_z, a.grad_bias, a.grad_scales = _ptr(gz_out), _ptr(gl), _ptr(gz), _ptr(gb), _ptr(gs)
    _call("cnf_actnorm_bwd", a, z, (z, b1, s1, pad, length, gz_out, gl))
    return gz, gb.reshape(bias.shape), gs.reshape(scales.shape)


def ext_actnorm_backward(z, ext, pad, z_out, g_z, g_ldj, reverse, needs):
    z = _f32(z, "z")
    B, S, Cc = z.shape
    ext = _f32(ext, "ext", (B, S, 2 * Cc))
    pad = _pad_bs(pad, B, S)
    gz_out = _grad(g_z, z)
    gl = _opt_f32(g_ldj, "grad_ldj", (B,))
    gz, gext = torch.empty_like(z), torch.empty_like(ext)
    a = L.ExtActnormBwdArgs()
    a.B, a.S, a.C = B, S, Cc
    a.z, a.ext, a.pad, a.reverse = _ptr(z), _ptr(ext), _ptr(pad), int(bool(reverse))
    a.grad_z_out, a.grad_ldj, a.grad_z, a.grad_ext = _ptr(gz_out), _ptr(gl), _ptr(gz), _ptr(gext)
    _call("cnf_ext_actnorm_bwd", a, z, (z, ext, pad, gz_out, gl))
    return gz, gext


def invconv_backward(z, weight, pad, length, g_z, g_ldj, reverse, needs):
    z = _f32(z, "z")
    B, S, Cc = z.shape
    weight = _f32(weight, "weight", (Cc, Cc))
    pad = _pad_bs(pad, B, S)
    length = _opt_f32(length, "length", (B,))
    gz_out = _grad(g_z, z)
    gl = _opt_f32(g_ldj, "grad_ldj", (B,))
    gz = torch.empty_like(z)
    gw = torch.zeros(Cc, Cc, dtype=torch.float32, device=z.device)
    gsl = torch.zeros(1, dtype=torch.float32, device=z.device)
    a = L.InvconvBwdArgs()
    a.B, a.S, a.C = B, S, Cc
    a.z, a.weight, a.pad, a.length, a.reverse = _ptr(z), _ptr(weight), _ptr(pad), _ptr(length), int(bool(reverse))
    a.grad_z_out, a.grad_ldj, a.grad_z, a.grad_weight, a.grad_sldj = _ptr(gz_out), _ptr(gl), _ptr(gz), _ptr(gw), _ptr(gsl)
    _call("cnf_invconv_bwd", a, z, (z, weight, pad, length, gz_out, gl))
    return gz, gw, gsl


def logistic_logprob_backward(x, g, mu, sigma):
    x = _f32(x, "x")
    shape = x.shape
    x3 = x.reshape(shape[0], -1, shape[-1]) if x.dim() >= 2 else x.reshape(1, 1, -1)
    g = _f32(g, "grad").reshape(x3.shape)
    gx = torch.empty_like(x3)
    a = L.LogisticLogprobBwdArgs()
    a.B, a.S, a.C = x3.shape
    a.x, a.mu, a.sigma = _ptr(x3), float(mu), float(sigma)
    a.grad_elementwise, a.grad_x = _ptr(g), _ptr(gx)
    _call("cnf_logistic_logprob_bwd", a, x3, (x3, g))
    return gx.reshape(shape)


def sigmoid_flow_backward(z, g_out, g_ldj, g_elem, reverse, alpha):
    z = _f32(z, "z")
    B = z.shape[0]
    gz = torch.empty_like(z)
    go = None if g_out is None else _f32(g_out, "grad_z_out")
    gl = _opt_f32(g_ldj, "grad_ldj", (B,))
    ge = None if g_elem is None else _f32(g_elem, "grad_ldj_elementwise")
    a = L.SigmoidFlowBwdArgs()
    a.B, a.n_per_sample, a.z, a.reverse, a.alpha = B, z.numel() // max(B, 1), _ptr(z), int(bool(reverse)), float(alpha)
    a.grad_z_out, a.grad_ldj, a.grad_ldj_elementwise, a.grad_z = _ptr(go), _ptr(gl), _ptr(ge), _ptr(gz)
    _call("cnf_sigmoid_flow_bwd", a, z, (z, go, gl, ge))
    return gz


def categ_encode_backward(tokens, z, table, category_prior, pad, beta, g_z, g_ldj):
    """dL/dtable [V,2D] of ``ops.categ_encode`` (``cnf_categ_encode_bwd``); ``z`` is the forward output."""
    tokens = tokens.long().contiguous()
    B, S = tokens.shape
    table = _f32(table, "table")
    V, D2 = table.shape
    D = D2 // 2
    z = _f32(z, "z", (B, S, D))
    prior = _f32(category_prior, "category_prior", (V,))
    pad = _pad_bs(pad, B, S)
    gz = _grad(g_z, z)
    gl = _opt_f32(g_ldj, "grad_ldj", (B,))
    gtable = torch.zeros_like(table)
    a = L.CategEncodeBwdArgs()
    a.B, a.S, a.V, a.D = B, S, V, D
    a.tokens, a.z, a.table, a.category_prior, a.pad, a.beta = _ptr(tokens), _ptr(z), _ptr(table), _ptr(prior), _ptr(pad), float(beta)
    a.grad_z, a.grad_ldj, a.grad_table = _ptr(gz), _ptr(gl), _ptr(gtable)
    _call("cnf_categ_encode_bwd", a, z, (tokens, z, table, prior, pad, gz, gl))
    return gtable
