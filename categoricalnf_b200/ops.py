"""Tensor-level entry points: torch CUDA tensors in, C-ABI calls out.

PyTorch is used here only as the owner of device memory and streams; every function launches
hand-written sm_100a kernels from ``libcnf_b200.so`` on ``torch.cuda.current_stream()``.  CPU
tensors are rejected - there is no fallback path.
"""
from __future__ import annotations

import ctypes as C
import weakref
import os
from typing import Optional, Sequence

import torch

from . import _lib as L

STRICT = os.environ.get("CNF_B200_STRICT", "0") not in ("", "0")

FLAG_NAN_Z, FLAG_NAN_LDJ, FLAG_CDF_RANGE = 1, 2, 4

_status_words = {}
_launches = 0  # number of C-ABI kernel launches issued by this process (bench.py reports it)


def launch_count() -> int:
    return _launches


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream(t: torch.Tensor) -> int:
    """cudaStream_t of torch's current stream on ``t``'s device (the raw-handle query is ~10x cheaper than building a
    ``torch.cuda.Stream`` object per launch, which matters for the launch-bound graph flows)."""
    if _raw_stream is not None:
        return _raw_stream(t.device.index)
    return torch.cuda.current_stream(t.device).cuda_stream


def _f32(t: torch.Tensor, name: str, shape=None) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor" % name)
    if not t.is_cuda:
        raise RuntimeError("categoricalnf_b200: %s lives on %s - the hot path runs on CUDA only "
                           "(there is no CPU fallback)" % (name, t.device))
    if t.dtype != torch.float32:
        t = t.float()
    if not t.is_contiguous():
        t = t.contiguous()
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise ValueError("%s has shape %s, expected %s" % (name, tuple(t.shape), tuple(shape)))
    return t


def _opt_f32(t, name, shape=None):
    return None if t is None else _f32(t, name, shape)


def _ldj(t: torch.Tensor, B: int) -> torch.Tensor:
    """A per-sample ldj that a kernel updates IN PLACE: it must already be a contiguous fp32 CUDA tensor of shape [B] -
    a silent ``.float()`` / ``.contiguous()`` copy would break the in-place contract (the caller's tensor would not change)."""
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError("categoricalnf_b200: ldj must be a CUDA tensor (there is no CPU fallback)")
    if t.dtype != torch.float32 or not t.is_contiguous() or tuple(t.shape) != (B,):
        raise ValueError("ldj is updated in place and must be a contiguous float32 tensor of shape [%d]; got %s %s%s"
                         % (B, t.dtype, tuple(t.shape), "" if t.is_contiguous() else " (non-contiguous)"))
    return t


def _pad_bs(pad, B, S, name="channel_padding_mask"):
    """Accept [B,S], [B,S,1] (or anything with B*S elements) and return a contiguous [B,S] view."""
    if pad is None:
        return None
    pad = _f32(pad, name)
    if pad.numel() != B * S:
        raise ValueError("%s has %d elements, expected B*S = %d" % (name, pad.numel(), B * S))
    return pad.reshape(B, S)


def _host_floats(vals: Optional[Sequence[float]]):
    if vals is None:
        return None, None
    arr = (C.c_float * len(vals))(*[float(v) for v in vals])
    return arr, C.cast(arr, C.c_void_p)


def _mask_struct(mask_c, mask_s):
    keep = []
    m = L.Mask()
    arr, p = _host_floats(mask_c)
    keep.append(arr)
    m.cond_c_host = p
    arr, p = _host_floats(mask_s)
    keep.append(arr)
    m.cond_s_host = p
    m.s_period = 0 if mask_s is None else len(mask_s)
    return m, keep


def status_word(device) -> torch.Tensor:
    """Per-device uint32 word that kernels OR health flags into (cnf_b200.h: CNF_FLAG_*)."""
    dev = torch.device(device)
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    w = _status_words.get(key)
    if w is None:
        w = torch.zeros(1, dtype=torch.int32, device=torch.device("cuda", key))
        _status_words[key] = w
    return w


def check_status(device, where: str = "") -> None:
    """Read and clear the status word (one host sync) and raise like the reference would:
    AssertionError for NaNs (mixture_cdf_layer.py:82, flow_model.py:42), RuntimeError for an
    inverse-CDF argument outside (0,1) (mixture_cdf_layer.py:238-239)."""
    w = status_word(device)
    bits = int(w.item())
    if bits == 0:
        return
    w.zero_()
    if bits & FLAG_CDF_RANGE:
        raise RuntimeError("Inverse logisitic CDF got y outside (0, 1)" + (" [%s]" % where if where else ""))
    what = []
    if bits & FLAG_NAN_Z:
        what.append("z")
    if bits & FLAG_NAN_LDJ:
        what.append("ldj")
    raise AssertionError("[!] ERROR: Found NaN in %s%s" % (" and ".join(what), " (%s)" % where if where else ""))


_cur_device = getattr(torch._C, "_cuda_getDevice", None) or torch.cuda.current_device
_set_device = getattr(torch._C, "_cuda_setDevice", None) or torch.cuda.set_device


def _call(name, args, ref: torch.Tensor, keep=None):
    """Launch ``name`` on torch's current stream of ``ref``'s device.  The C side launches on the calling thread's current
    device, so that device is switched to ``ref``'s for the duration of the call when it differs (model on cuda:1 while
    cuda:0 is current); every tensor in ``keep`` must live on the same device."""
    global _launches
    dev = ref.device.index
    if keep is not None:
        _same_device(dev, keep, name)
    cur = _cur_device()
    if cur != dev:
        _set_device(dev)
        try:
            L.call(name, args, _stream(ref))
        finally:
            _set_device(cur)
    else:
        L.call(name, args, _stream(ref))
    _launches += 1
    if STRICT:
        check_status(ref.device, name)


def _same_device(dev, items, name):
    for t in items:
        if isinstance(t, torch.Tensor):
            if t.is_cuda and t.device.index != dev:
                raise RuntimeError("categoricalnf_b200: %s got tensors on cuda:%d and cuda:%d - all arguments of one call "
                                   "must live on one device" % (name, dev, t.device.index))
        elif isinstance(t, (tuple, list)):
            _same_device(dev, t, name)


# ----------------------------------------------------------------------------------------------
def _n_transformed(mask_c, Cc):
    return Cc if mask_c is None else sum(1 for m in mask_c if float(m) == 0.0)


def _mixcdf_args(z, nn_out, num_mixtures, mask_c, mask_s, pad, scaling_factor, mixture_scaling_factor, compact=False):
    z = _f32(z, "z")
    if z.dim() != 3:
        raise ValueError("z must be [B, S, C]")
    B, S, Cc = z.shape
    K = int(num_mixtures)
    nn_out = _f32(nn_out, "nn_out", (B, S, (_n_transformed(mask_c, Cc) if compact else Cc) * (2 + 3 * K)))
    pad = _pad_bs(pad, B, S)
    a = L.MixcdfArgs()
    a.B, a.S, a.C, a.K = B, S, Cc, K
    a.mask, keep = _mask_struct(mask_c, mask_s)
    sf = _opt_f32(scaling_factor, "scaling_factor", (Cc,))
    msf = _opt_f32(mixture_scaling_factor, "mixture_scaling_factor", (Cc, K))
    a.z, a.nn_out, a.pad = _ptr(z), _ptr(nn_out), _ptr(pad)
    a.scaling_factor, a.mixture_scaling_factor = _ptr(sf), _ptr(msf)
    a.nn_compact = int(bool(compact))
    return a, (keep, z, nn_out, pad, sf, msf)


def mixcdf_fusable(z, nn_out, num_mixtures, *, mask_c=None, mask_s=None, prebounded=False):
    """True when ``mixcdf(..., fuse_next=...)`` is available for this shape / mask / alignment."""
    a, keep = _mixcdf_args(z, nn_out, num_mixtures, mask_c, mask_s, None, None, None)
    a.params_prebounded = int(bool(prebounded))
    return bool(L.load().cnf_mixcdf_fusable(C.byref(a)))


MIXCDF_PATHS = {0: "generic", 1: "pipe", 2: "gpipe"}


def mixcdf_path(z, nn_out, num_mixtures, *, mask_c=None, mask_s=None, prebounded=False, compact=False):
    """Which kernel :func:`mixcdf` launches for this shape / mask / alignment: "generic" (staged), "pipe" (TMA pipeline,
    compile-time K and Ct, thread per element) or "gpipe" (TMA pipeline, lane groups, any K).  ``compact``: ``nn_out`` holds
    only the transformed channels' records ([B,S,Ct*(2+3K)]) - taken by the two pipelines only."""
    a, keep = _mixcdf_args(z, nn_out, num_mixtures, mask_c, mask_s, None, None, None, compact)
    a.params_prebounded = int(bool(prebounded))
    return MIXCDF_PATHS[int(L.load().cnf_mixcdf_path(C.byref(a)))]


def mixcdf(z, nn_out, num_mixtures, *, mask_c=None, mask_s=None, pad=None, scaling_factor=None,
           mixture_scaling_factor=None, reverse=False, reg_max=-1.0, reg_factor=1.0, training=False,
           ldj=None, want_reg=False, out=None, prebounded=False, fuse_next=None, compact=False):
    """K1/K2.  Returns ``(z_out, ldj[B], reg_ldj[B] | None)``.  ``ldj`` given -> accumulated into.
    ``compact``: ``nn_out`` is [B,S,Ct*(2+3K)] - the records of the transformed channels only (one contiguous run); the
    conditioner half of the network output is never read, so its producer need not compute it (see ``mixcdf_path``).
    ``scaling_factor`` / ``mixture_scaling_factor`` None = zeros, i.e. tanh bounds e^0 = 1 like a freshly built layer (the
    module always passes its parameters; the reference's static helper treats None as "no bounding", which is
    ``prebounded=True`` here).
    ``fuse_next = (bias [C], scales [C], W [C,C])`` applies the next block's ActNorm and 1x1
    convolution to the output row inside the kernel (forward only, see ``mixcdf_fusable``); their
    per-sample-constant ldj terms are NOT added here."""
    a, keep = _mixcdf_args(z, nn_out, num_mixtures, mask_c, mask_s, pad, scaling_factor, mixture_scaling_factor, compact)
    z = keep[1]
    B = z.shape[0]
    z_out = torch.empty_like(z) if out is None else out
    accumulate = ldj is not None
    ldj_t = _ldj(ldj, B) if accumulate else torch.empty(B, dtype=torch.float32, device=z.device)
    reg = torch.empty(B, dtype=torch.float32, device=z.device) if want_reg else None
    a.reg_max, a.reg_factor, a.training = float(reg_max), float(reg_factor), int(bool(training))
    a.accumulate = int(accumulate)
    a.params_prebounded = int(bool(prebounded))
    a.z_out, a.ldj, a.reg_ldj = _ptr(z_out), _ptr(ldj_t), _ptr(reg)
    a.status = _ptr(status_word(z.device))
    if fuse_next is not None:
        Cc = z.shape[2]
        nb = _f32(fuse_next[0], "next bias").reshape(-1)
        ns = _f32(fuse_next[1], "next scales").reshape(-1)
        nw = _f32(fuse_next[2], "next conv weight", (Cc, Cc))
        keep = keep + (nb, ns, nw)
        a.next_actnorm_bias, a.next_actnorm_scales, a.next_conv_weight = _ptr(nb), _ptr(ns), _ptr(nw)
    _call("cnf_mixcdf_inv" if reverse else "cnf_mixcdf_fwd", a, z, keep)
    return z_out, ldj_t, reg


def affine_coupling(z, nn_out, ldj, *, mask_c=None, mask_s=None, scaling_factor=None, reverse=False,
                    prebounded=False):
    """K3.  ``ldj`` [B] is updated in place and returned together with z_out."""
    z = _f32(z, "z")
    B, S, Cc = z.shape
    nn_out = _f32(nn_out, "nn_out", (B, S, 2 * Cc))
    ldj = _ldj(ldj, B)
    a = L.AffineArgs()
    a.B, a.S, a.C = B, S, Cc
    a.mask, keep = _mask_struct(mask_c, mask_s)
    sf = _opt_f32(scaling_factor, "scaling_factor", (Cc,))
    z_out = torch.empty_like(z)
    a.z, a.nn_out, a.scaling_factor, a.reverse = _ptr(z), _ptr(nn_out), _ptr(sf), int(bool(reverse))
    a.params_prebounded = int(bool(prebounded))
    a.z_out, a.ldj, a.status = _ptr(z_out), _ptr(ldj), _ptr(status_word(z.device))
    _call("cnf_affine_coupling", a, z, keep)
    return z_out, ldj


def actnorm(z, bias, scales, ldj=None, *, pad=None, length=None, reverse=False):
    """K4.  ``ldj`` (if given) is updated IN PLACE like the reference's ``ldj +=``."""
    z = _f32(z, "z")
    B, S, Cc = z.shape
    a = L.ActnormArgs()
    a.B, a.S, a.C = B, S, Cc
    bias = _f32(bias, "bias").reshape(-1)
    scales = _f32(scales, "scales").reshape(-1)
    pad = _pad_bs(pad, B, S)
    length = _opt_f32(length, "length", (B,))
    if ldj is not None:
        ldj = _ldj(ldj, B)
    z_out = torch.empty_like(z)
    a.z, a.bias, a.scales, a.pad, a.length = _ptr(z), _ptr(bias), _ptr(scales), _ptr(pad), _ptr(length)
    a.reverse, a.z_out, a.ldj, a.status = int(bool(reverse)), _ptr(z_out), _ptr(ldj), _ptr(status_word(z.device))
    _call("cnf_actnorm", a, z)
    return z_out, ldj


def ext_actnorm(z, ext, ldj, *, pad=None, reverse=False):
    """K4 (external).  ``ext`` = pred_net output ``[B,S,2C]``; ``ldj`` updated in place."""
    z = _f32(z, "z")
    B, S, Cc = z.shape
    ext = _f32(ext, "ext", (B, S, 2 * Cc))
    ldj = _ldj(ldj, B)
    pad = _pad_bs(pad, B, S)
    a = L.ExtActnormArgs()
    a.B, a.S, a.C = B, S, Cc
    z_out = torch.empty_like(z)
    a.z, a.ext, a.pad, a.reverse = _ptr(z), _ptr(ext), _ptr(pad), int(bool(reverse))
    a.z_out, a.ldj, a.status = _ptr(z_out), _ptr(ldj), _ptr(status_word(z.device))
    _call("cnf_ext_actnorm", a, z)
    return z_out, ldj


def actnorm_data_init(x, pad=None):
    """Masked per-channel statistics -> (bias [C], scales [C]) (activation_normalization.py:55-67)."""
    x = _f32(x, "x")
    B, S, Cc = x.shape
    pad = _pad_bs(pad, B, S)
    ws = torch.empty(3 * Cc, dtype=torch.float64, device=x.device)
    bias = torch.empty(Cc, dtype=torch.float32, device=x.device)
    scales = torch.empty(Cc, dtype=torch.float32, device=x.device)
    a = L.ActnormInitArgs()
    a.B, a.S, a.C = B, S, Cc
    a.x, a.pad, a.workspace, a.bias, a.scales = _ptr(x), _ptr(pad), _ptr(ws), _ptr(bias), _ptr(scales)
    _call("cnf_actnorm_data_init", a, x)
    return bias, scales


def invconv_build(p=None, l=None, u=None, log_s=None, sign_s=None, weight=None, want_inverse=True):
    """K5 build.  Returns ``(W [C,C], W_inv [C,C] | None, sldj [1])`` as device tensors."""
    ref = weight if weight is not None else l
    ref = _f32(ref, "weight/l")
    Cc = ref.shape[0]
    a = L.InvconvBuildArgs()
    a.C = Cc
    ts = {k: _opt_f32(v, k) for k, v in dict(p=p, l=l, u=u, log_s=log_s, sign_s=sign_s, weight=weight).items()}
    w = torch.empty(Cc, Cc, dtype=torch.float32, device=ref.device)
    w_inv = torch.empty(Cc, Cc, dtype=torch.float32, device=ref.device) if want_inverse else None
    sldj = torch.empty(1, dtype=torch.float32, device=ref.device)
    a.p, a.l, a.u, a.log_s, a.sign_s, a.weight = (_ptr(ts[k]) for k in ("p", "l", "u", "log_s", "sign_s", "weight"))
    a.w_out, a.w_inv_out, a.sldj_out = _ptr(w), _ptr(w_inv), _ptr(sldj)
    _call("cnf_invconv_build", a, ref)
    return w, w_inv, sldj


def invconv_apply(z, weight, sldj, ldj=None, *, pad=None, length=None, reverse=False, pre_actnorm=None, out_mask=None):
    """K5 apply.  Returns ``(z_out, ldj)``; ldj (if given) is updated in place.

    ``pre_actnorm = (bias [C], scales [C])`` applies the block's ActNorm first in the same pass (forward only;
    its per-sample-constant ldj term is NOT added here).  With ``out_mask`` [C] a third value
    ``z_out * out_mask`` is returned (the masked network input of the following coupling layer)."""
    z = _f32(z, "z")
    B, S, Cc = z.shape
    weight = _f32(weight, "weight", (Cc, Cc))
    sldj = _f32(sldj, "sldj").reshape(1)
    pad = _pad_bs(pad, B, S)
    length = _opt_f32(length, "length", (B,))
    if ldj is not None:
        ldj = _ldj(ldj, B)
    z_out = torch.empty_like(z)
    a = L.InvconvArgs()
    a.B, a.S, a.C = B, S, Cc
    a.z, a.weight, a.sldj, a.pad, a.length = _ptr(z), _ptr(weight), _ptr(sldj), _ptr(pad), _ptr(length)
    a.reverse, a.z_out, a.ldj, a.status = int(bool(reverse)), _ptr(z_out), _ptr(ldj), _ptr(status_word(z.device))
    keep, z_masked = (z, weight, sldj, pad, length), None
    if pre_actnorm is not None:
        pb = _f32(pre_actnorm[0], "actnorm bias").reshape(-1)
        ps = _f32(pre_actnorm[1], "actnorm scales").reshape(-1)
        keep = keep + (pb, ps)
        a.pre_actnorm_bias, a.pre_actnorm_scales = _ptr(pb), _ptr(ps)
    if out_mask is not None:
        om = _f32(out_mask, "out_mask").reshape(-1)
        if om.numel() != Cc:
            raise ValueError("out_mask has %d entries, expected %d" % (om.numel(), Cc))
        z_masked = torch.empty_like(z)
        keep = keep + (om,)
        a.out_mask, a.z_masked_out = _ptr(om), _ptr(z_masked)
    _call("cnf_invconv_apply", a, z, keep)
    if z_masked is not None:
        return z_out, ldj, z_masked
    return z_out, ldj


def categ_encode(tokens, table, category_prior, ldj, *, noise=None, seed=0, offset=0, pad=None, beta=1.0,
                 want_class_prob=False, fuse_next=None):
    """K6 encode.  tokens [B,S] int64 -> (z [B,S,D], ldj (in place), class_prob_log [B,S] | None).
    ``fuse_next = (bias [D], scales [D], W [D,D])`` applies the first block's ActNorm + 1x1 convolution
    inside the kernel when the shape allows it; returns None as first element otherwise is NOT done -
    call :func:`categ_encode_fusable` first."""
    if not tokens.is_cuda:
        raise RuntimeError("categoricalnf_b200: tokens live on %s - CUDA only" % tokens.device)
    tokens = tokens.long().contiguous()
    B, S = tokens.shape
    table = _f32(table, "table")
    V, D2 = table.shape
    D = D2 // 2
    prior = _opt_f32(category_prior, "category_prior", (V,))
    ldj = _ldj(ldj, B)
    pad = _pad_bs(pad, B, S)
    if noise is not None:
        noise = _f32(noise, "noise")
        if noise.numel() != B * S * D:
            raise ValueError("noise has %d elements, expected %d" % (noise.numel(), B * S * D))
    z = torch.empty(B, S, D, dtype=torch.float32, device=tokens.device)
    cpl = torch.empty(B, S, dtype=torch.float32, device=tokens.device) if want_class_prob else None
    a = L.CategEncodeArgs()
    a.B, a.S, a.V, a.D = B, S, V, D
    a.tokens, a.u_noise, a.seed, a.offset = _ptr(tokens), _ptr(noise), int(seed), int(offset)
    a.table, a.category_prior, a.pad, a.beta = _ptr(table), _ptr(prior), _ptr(pad), float(beta)
    a.z_out, a.ldj, a.class_prob_log, a.status = _ptr(z), _ptr(ldj), _ptr(cpl), _ptr(status_word(z.device))
    keep = None
    if fuse_next is not None:
        nb = _f32(fuse_next[0], "next bias").reshape(-1)
        ns = _f32(fuse_next[1], "next scales").reshape(-1)
        nw = _f32(fuse_next[2], "next conv weight", (D, D))
        keep = (nb, ns, nw)
        a.next_actnorm_bias, a.next_actnorm_scales, a.next_conv_weight = _ptr(nb), _ptr(ns), _ptr(nw)
    _call("cnf_categ_encode", a, z, keep)
    return z, ldj, cpl


def categ_encode_fusable(B, S, V, D):
    """True when ``categ_encode(..., fuse_next=...)`` is available for this problem size."""
    a = L.CategEncodeArgs()
    a.B, a.S, a.V, a.D = B, S, V, D
    return bool(L.load().cnf_categ_encode_fusable(C.byref(a)))


def categ_decode(z, table, category_prior):
    """K6 decode.  z [B,S,D] -> tokens [B,S] int64 (argmax of the class-conditional density)."""
    z = _f32(z, "z")
    B, S, D = z.shape
    table = _f32(table, "table")
    V = table.shape[0]
    if table.shape[1] != 2 * D:
        raise ValueError("table is %s but z has D=%d" % (tuple(table.shape), D))
    prior = _opt_f32(category_prior, "category_prior", (V,))
    out = torch.empty(B, S, dtype=torch.int64, device=z.device)
    a = L.CategDecodeArgs()
    a.B, a.S, a.V, a.D = B, S, V, D
    a.z, a.table, a.category_prior, a.tokens_out = _ptr(z), _ptr(table), _ptr(prior), _ptr(out)
    _call("cnf_categ_decode", a, z)
    return out


def logistic_logprob(x, *, pad=None, mu=0.0, sigma=1.0 / 1.81, reduce=True, elementwise=False, out=None, add=None, total=None):
    """K7.  Returns per-sample sums [B] (``reduce``) and/or the element-wise log-density.

    ``add`` [B] (e.g. the flow's ldj): the returned per-sample values are ``add + log_prob`` - the per-sample
    log-likelihood - written by the same kernel.  ``total`` (float64 [2], CUDA): receives ``(sum_b result[b], B)``, the pair
    the ranks all-reduce once per step (``sharding.LogLikAllReducer``), from the kernel's epilogue - no separate reduction."""
    x = _f32(x, "x")
    shape = x.shape
    x3 = x.reshape(shape[0], -1, shape[-1]) if x.dim() >= 2 else x.reshape(1, 1, -1)
    B, S, Cc = x3.shape
    pad = _pad_bs(pad, B, S) if pad is not None else None
    a = L.LogisticLogprobArgs()
    a.B, a.S, a.C = B, S, Cc
    acc = out is not None
    if (add is not None or total is not None) and (acc or not reduce):
        raise ValueError("add / total need reduce=True and no `out` to accumulate into")
    res = _ldj(out, B) if acc else (torch.empty(B, dtype=torch.float32, device=x.device) if reduce else None)
    elem = torch.empty_like(x) if elementwise else None
    add = _opt_f32(add, "add", (B,))
    if total is not None and (not total.is_cuda or total.dtype != torch.float64 or total.numel() != 2 or not total.is_contiguous()):
        raise ValueError("total must be a contiguous CUDA float64 tensor with 2 elements")
    a.x, a.pad, a.mu, a.sigma, a.accumulate = _ptr(x3), _ptr(pad), float(mu), float(sigma), int(acc)
    a.out, a.elementwise, a.add, a.total = _ptr(res), _ptr(elem), _ptr(add), _ptr(total)
    _call("cnf_logistic_logprob", a, x, (x3, pad, add, total))
    return res, elem


def logistic_sample(shape, device, *, noise=None, seed=0, offset=0, mu=0.0, sigma=1.0 / 1.81, eps=1e-4):
    """K7 sampler: logit of a squeezed uniform, scaled (distributions.py:139-145)."""
    x = torch.empty(tuple(shape), dtype=torch.float32, device=device)
    if not x.is_cuda:
        raise RuntimeError("categoricalnf_b200: sampling runs on CUDA only")
    if noise is not None:
        noise = _f32(noise, "noise")
    a = L.LogisticSampleArgs()
    a.n, a.u_noise, a.seed, a.offset = x.numel(), _ptr(noise), int(seed), int(offset)
    a.mu, a.sigma, a.eps, a.x_out = float(mu), float(sigma), float(eps), _ptr(x)
    _call("cnf_logistic_sample", a, x)
    return x


def sigmoid_flow(z, ldj=None, *, reverse=False, alpha=1e-5, sum_ldj=True, add_tokens=None):
    """SigmoidFlow (sigmoid_layer.py:24-48) in its EFFECTIVE direction: ``reverse=False`` sigmoid, ``True`` logit of the
    alpha-squeezed input.  Returns ``(z_out, ldj)``: ``ldj`` is a NEW tensor ``ldj_in + sum`` [B] (``sum_ldj``) or the
    element-wise values.  ``add_tokens`` (int64, z's shape) is added to ``z_out`` (variational_dequantization.py:47)."""
    z = _f32(z, "z")
    B = z.shape[0] if z.dim() >= 1 else 1
    per = z.numel() // max(B, 1)
    a = L.SigmoidFlowArgs()
    a.B, a.n_per_sample, a.z, a.reverse, a.alpha = B, per, _ptr(z), int(bool(reverse)), float(alpha)
    out = torch.empty_like(z)
    elem = None
    if sum_ldj:
        res = torch.zeros(B, dtype=torch.float32, device=z.device) if ldj is None else _f32(ldj, "ldj", (B,)).clone()
        a.accumulate, a.ldj = 1, _ptr(res)
    else:
        elem = torch.empty_like(z)
        a.accumulate, a.ldj_elementwise = 0, _ptr(elem)
    if add_tokens is not None:
        add_tokens = add_tokens.long().contiguous()
        if add_tokens.numel() != z.numel() or not add_tokens.is_cuda:
            raise ValueError("add_tokens must be a CUDA int64 tensor with as many elements as z")
        a.add_tokens = _ptr(add_tokens)
    a.z_out, a.status = _ptr(out), _ptr(status_word(z.device))
    _call("cnf_sigmoid_flow", a, z, (add_tokens,))
    return out, (res if sum_ldj else elem)


def dequant_floor(z, vocab_size):
    """tokens = clamp(floor(z), 0, V-1) as int64 (variational_dequantization.py:55-56); a trailing singleton dim is dropped."""
    z = _f32(z, "z")
    out = torch.empty(z.shape[:-1] if z.dim() > 2 and z.shape[-1] == 1 else z.shape, dtype=torch.int64, device=z.device)
    a = L.DequantFloorArgs()
    a.n, a.V, a.z, a.tokens_out = z.numel(), int(vocab_size), _ptr(z), _ptr(out)
    _call("cnf_dequant_floor", a, z)
    return out


def ldj_axpy(y, *, alpha=1.0, alpha_dev=None, x=None, length=None):
    """y[b] += alpha * alpha_dev * x[b] * length[b] (missing factors are 1)."""
    y = _f32(y, "y")
    a = L.LdjAxpyArgs()
    a.B, a.alpha = y.numel(), float(alpha)
    a.alpha_dev, a.x, a.length, a.y = _ptr(_opt_f32(alpha_dev, "alpha_dev")), _ptr(_opt_f32(x, "x")), \
        _ptr(_opt_f32(length, "length")), _ptr(y)
    _call("cnf_ldj_axpy", a, y)
    return y


# ----------------------------------------------------------------------------------------------
# K8: dense projections of the coupling networks on tcgen05 tensor cores
# ----------------------------------------------------------------------------------------------
PRECISION = {"tf32": 0, "3xtf32": 1}
ACTIVATION = {None: 0, "none": 0, "gelu": 1}


linear_profile = None        # set to a list: every cnf_linear_fwd launch appends (M, N, K, precision, start event, end event)
_weight_split_cache = {}     # id(tensor) -> (weakref, key, hi, lo)
_param_epoch = 0             # bumped whenever parameters may have changed without their ``_version`` moving


def param_epoch() -> int:
    """Generation counter of everything DERIVED from parameters (3xTF32 weight splits, fused projection weights, built 1x1
    convolution matrices, captured CUDA graphs).  ``tensor._version`` alone cannot key those caches: in-place writes through
    ``p.data`` - which is how the reference's default optimiser updates weights (general/radam.py:82,147) - leave it
    unchanged.  The counter moves on every optimiser step (a global ``register_optimizer_step_post_hook``), on every
    ``train()`` / ``eval()`` switch of a drop-in module and on :func:`invalidate_caches`; and caches are bypassed
    altogether while gradients are being recorded."""
    return _param_epoch


def invalidate_caches(*_args, **_kwargs) -> None:
    """Call after changing parameters in a way autograd's version counters cannot see (``p.data`` writes outside an
    ``torch.optim.Optimizer.step``, e.g. swapping in EMA weights)."""
    global _param_epoch
    _param_epoch += 1


def param_fingerprint(tensors) -> tuple:
    """Cheap identity of a set of parameters / buffers for caches that hold DERIVED device tensors (captured CUDA graphs):
    the cache generation, the version counters and the storage addresses."""
    v = a = 0
    for t in tensors:
        v += t._version
        a ^= t.data_ptr()
    return (_param_epoch, v, a)


try:        # every torch.optim.Optimizer subclass - the reference's RAdam / Adam included - runs this after step()
    from torch.optim.optimizer import register_optimizer_step_post_hook as _reg_post_step
    _reg_post_step(invalidate_caches)
except ImportError:      # pragma: no cover - torch < 2.0
    pass


def _rna_tf32(t):
    """Round fp32 to the nearest TF32 value (10 mantissa bits), ties away from zero - ``cvt.rna.tf32.f32``."""
    return ((t.contiguous().view(torch.int32) + 0x1000) & -8192).view(torch.float32)


def weight_split(weight, use_cache=True):
    """(hi, lo) with hi = rna_tf32(weight), lo = rna_tf32(weight - hi): the 3xTF32 split of a weight, done on the host
    side of the ABI instead of once per tile inside the kernel.  Only long-lived weights are split here: ``nn.Parameter``s
    and tensors marked ``_cnf_cache_lo`` (fused weight blocks); None for anything else.  ``use_cache=False`` (training:
    the weight changes every step) splits afresh and stores nothing; otherwise the split is cached per tensor object,
    version and :func:`param_epoch`."""
    if not (isinstance(weight, torch.nn.Parameter) or getattr(weight, "_cnf_cache_lo", False)):
        return None
    if not use_cache or (torch.is_grad_enabled() and weight.requires_grad):
        _weight_split_cache.pop(id(weight), None)
        with torch.no_grad():
            w = weight.detach()
            hi = _rna_tf32(w)
            return hi, _rna_tf32(w - hi)
    key = (weight._version, weight.data_ptr(), tuple(weight.shape), _param_epoch)
    hit = _weight_split_cache.get(id(weight))
    if hit is None or hit[0]() is not weight or hit[1] != key:
        with torch.no_grad():
            w = weight.detach()
            hi = _rna_tf32(w)
            lo = _rna_tf32(w - hi)
        wid = id(weight)
        ref = weakref.ref(weight, lambda _r, wid=wid: _weight_split_cache.pop(wid, None))
        hit = (ref, key, hi, lo)
        _weight_split_cache[wid] = hit
    return hit[2], hit[3]


def linear(x, weight, bias=None, *, precision="3xtf32", activation=None, block_n=0, cache_weight=True, split=None):
    """y = x @ weight.T + bias (nn.Linear) on the tensor cores (``cnf_linear_fwd``).

    ``x`` [..., K] fp32 CUDA, ``weight`` [N, K], ``bias`` [N] | None.  ``precision``: "tf32" (one pass)
    or "3xtf32" (hi/lo split, fp32-level accuracy - default, keeps the 1e-4 parity of the flow).
    K that is not a multiple of 4 is zero-padded (a copy); everything else runs in place.  ``cache_weight=False``: the
    3xTF32 split of the weight is recomputed for this call (training step: the weight is about to change).  ``split`` =
    (hi, lo): a split the caller made already with :func:`weight_split` (it keeps it for the backward pass)."""
    x = _f32(x, "x")
    if split is None and precision == "3xtf32" and weight.dtype == torch.float32 and weight.is_contiguous():
        split = weight_split(weight, cache_weight)
    weight = _f32(weight, "weight")
    w_lo = None
    if split is not None:
        weight, w_lo = split          # B operand = exactly representable high part, low part from the cache
    N, K = weight.shape
    if x.shape[-1] != K:
        raise ValueError("x has %d input features, weight expects %d" % (x.shape[-1], K))
    lead = x.shape[:-1]
    x2 = x.reshape(-1, K)
    if K % 4 != 0:
        padk = 4 - K % 4
        x2 = torch.nn.functional.pad(x2, (0, padk))
        weight = torch.nn.functional.pad(weight, (0, padk))
        w_lo = torch.nn.functional.pad(w_lo, (0, padk)) if w_lo is not None else None
        K += padk
    if x2.data_ptr() % 16 != 0:
        x2 = x2.clone()
    bias = _opt_f32(bias, "bias", (N,))
    y = torch.empty(x2.shape[0], N, dtype=torch.float32, device=x.device)
    a = L.LinearArgs()
    a.M, a.N, a.K = x2.shape[0], N, K
    a.x, a.weight, a.bias, a.y = _ptr(x2), _ptr(weight), _ptr(bias), _ptr(y)
    a.precision, a.activation = PRECISION[precision], ACTIVATION[activation]
    a.weight_lo, a.block_n = _ptr(w_lo), int(block_n)
    if linear_profile is None:
        _call("cnf_linear_fwd", a, x2, (x2, weight, bias, w_lo))
    else:       # bench.py's tensor-pipe roofline: CUDA events around every projection launch on the launching stream
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _call("cnf_linear_fwd", a, x2, (x2, weight, bias, w_lo))
        e1.record()
        linear_profile.append((a.M, N, K, precision, e0, e1))
    return y.reshape(lead + (N,))


_BWD_NO_PRESPLIT = bool(os.environ.get("CNF_B200_BWD_NO_PRESPLIT"))      # A/B switch: grad_x splits the weight tile in the kernel


def linear_bwd(x, weight, grad_y, *, need_x=True, need_weight=True, need_bias=False, precision="3xtf32",
               grad_weight=None, grad_bias=None, weight_split=None):
    """Backward of :func:`linear` (``cnf_linear_bwd``): ``(grad_x | None, grad_weight | None, grad_bias | None)``.

    The three products read ``x`` / ``weight`` / ``grad_y`` in place (MN-major tensor-core operands, no transposed
    copies).  ``grad_weight`` / ``grad_bias`` given -> accumulated into (gradient accumulation); otherwise fresh
    zero-initialised tensors.  N or K not a multiple of 4 is zero-padded (copies).  ``weight_split = (hi, lo)`` (3xTF32): the
    split of ``weight`` the forward pass made (:func:`weight_split`) - the grad_x product then reads it through TMA instead of
    splitting the weight tile in every CTA and k-block."""
    x = _f32(x, "x")
    weight = _f32(weight, "weight")
    grad_y = _f32(grad_y, "grad_y")
    N, K = weight.shape
    x2 = x.reshape(-1, K)
    gy2 = grad_y.reshape(-1, N)
    M = x2.shape[0]
    if gy2.shape[0] != M:
        raise ValueError("grad_y has %d rows, x has %d" % (gy2.shape[0], M))
    padn, padk = (-N) % 4, (-K) % 4
    if padn or padk:       # rare (graph nets use multiples of 4): pad, run, slice
        F = torch.nn.functional
        gx, gw, gb = linear_bwd(F.pad(x2, (0, padk)), F.pad(weight, (0, padk, 0, padn)), F.pad(gy2, (0, padn)),
                                need_x=need_x, need_weight=need_weight, need_bias=need_bias, precision=precision)
        gx = gx[:, :K].reshape(x.shape) if gx is not None else None
        gw = gw[:N, :K] if gw is not None else None
        gb = gb[:N] if gb is not None else None
        if gw is not None and grad_weight is not None:
            gw = grad_weight.add_(gw)
        if gb is not None and grad_bias is not None:
            gb = grad_bias.add_(gb)
        return gx, gw, gb
    if x2.data_ptr() % 16 != 0:
        x2 = x2.clone()
    if gy2.data_ptr() % 16 != 0:
        gy2 = gy2.clone()
    gx = torch.empty(M, K, dtype=torch.float32, device=x.device) if need_x else None
    gw = gb = None
    if need_weight:
        gw = _f32(grad_weight, "grad_weight", (N, K)) if grad_weight is not None else \
            torch.zeros(N, K, dtype=torch.float32, device=x.device)
    if need_bias:
        gb = _f32(grad_bias, "grad_bias", (N,)) if grad_bias is not None else torch.zeros(N, dtype=torch.float32, device=x.device)
    a = L.LinearBwdArgs()
    a.M, a.N, a.K = M, N, K
    a.x, a.weight, a.grad_y = _ptr(x2), _ptr(weight), _ptr(gy2)
    if weight_split is not None and precision == "3xtf32" and need_x and tuple(weight_split[0].shape) == (N, K) and not _BWD_NO_PRESPLIT:
        w_hi, w_lo = _f32(weight_split[0], "weight (high part)"), _f32(weight_split[1], "weight (low part)")
        a.weight, a.weight_lo = _ptr(w_hi), _ptr(w_lo)
    else:
        w_hi = w_lo = None
    a.precision = PRECISION[precision]
    a.grad_x, a.grad_weight, a.grad_bias = _ptr(gx), _ptr(gw), _ptr(gb)
    _call("cnf_linear_bwd", a, x2, (x2, weight, gy2, gx, gw, gb, w_hi, w_lo))
    return (gx.reshape(x.shape) if gx is not None else None), gw, gb


def _linear_mixcdf_args(z, features, weight, bias, num_mixtures, mask_c, mask_s, pad, scaling_factor,
                        mixture_scaling_factor, precision):
    z = _f32(z, "z")
    if z.dim() != 3:
        raise ValueError("z must be [B, S, C]")
    B, S, Cc = z.shape
    K = int(num_mixtures)
    features = _f32(features, "features")
    H = features.shape[-1]
    if features.numel() != B * S * H:
        raise ValueError("features has shape %s, expected [%d, %d, H]" % (tuple(features.shape), B, S))
    weight = _f32(weight, "weight", (Cc * (2 + 3 * K), H))
    bias = _opt_f32(bias, "bias", (Cc * (2 + 3 * K),))
    pad = _pad_bs(pad, B, S)
    a = L.LinearMixcdfArgs()
    a.mix.B, a.mix.S, a.mix.C, a.mix.K = B, S, Cc, K
    a.mix.mask, keep = _mask_struct(mask_c, mask_s)
    sf = _opt_f32(scaling_factor, "scaling_factor", (Cc,))
    msf = _opt_f32(mixture_scaling_factor, "mixture_scaling_factor", (Cc, K))
    a.mix.z, a.mix.pad = _ptr(z), _ptr(pad)
    a.mix.scaling_factor, a.mix.mixture_scaling_factor = _ptr(sf), _ptr(msf)
    a.H, a.precision = H, PRECISION[precision]
    a.features, a.weight, a.bias = _ptr(features), _ptr(weight), _ptr(bias)
    return a, (keep, z, features, weight, bias, pad, sf, msf)


def linear_mixcdf_fusable(z, features, weight, num_mixtures, *, mask_c=None, mask_s=None):
    """True when :func:`linear_mixcdf` can run this shape / mask / alignment in one kernel."""
    a, keep = _linear_mixcdf_args(z, features, weight, None, num_mixtures, mask_c, mask_s, None, None, None, "3xtf32")
    return bool(L.load().cnf_linear_mixcdf_fusable(C.byref(a)))


def linear_mixcdf(z, features, weight, bias, num_mixtures, *, mask_c=None, mask_s=None, pad=None, scaling_factor=None,
                  mixture_scaling_factor=None, reverse=False, reg_max=-1.0, reg_factor=1.0, training=False, ldj=None,
                  want_reg=False, precision="3xtf32", fuse_next=None, next_mask=None):
    """Final projection of the coupling network + mixture coupling transform in ONE kernel:
    ``mixcdf(z, features @ weight.T + bias, ...)`` without materialising the network output
    (``cnf_linear_mixcdf_fwd`` / ``_inv``).  Returns ``(z_out, ldj [B], reg_ldj [B] | None)``.

    ``fuse_next = (bias [C], scales [C], W [C,C])`` additionally applies the next block's ActNorm + 1x1
    convolution to the finished row (forward only; their per-sample-constant ldj terms are NOT added here).
    With ``next_mask`` ([C], the next coupling's mask) a fourth value ``z_out * next_mask`` is returned - the
    network input of that coupling (coupling_layer.py:53) - saving its separate masking pass."""
    a, keep = _linear_mixcdf_args(z, features, weight, bias, num_mixtures, mask_c, mask_s, pad, scaling_factor,
                                  mixture_scaling_factor, precision)
    z = keep[1]
    B = z.shape[0]
    z_out = torch.empty_like(z)
    accumulate = ldj is not None
    ldj_t = _ldj(ldj, B) if accumulate else torch.empty(B, dtype=torch.float32, device=z.device)
    reg = torch.empty(B, dtype=torch.float32, device=z.device) if want_reg else None
    a.mix.reg_max, a.mix.reg_factor, a.mix.training = float(reg_max), float(reg_factor), int(bool(training))
    a.mix.accumulate = int(accumulate)
    a.mix.z_out, a.mix.ldj, a.mix.reg_ldj = _ptr(z_out), _ptr(ldj_t), _ptr(reg)
    a.mix.status = _ptr(status_word(z.device))
    z_masked = None
    if fuse_next is not None:
        Cc = z.shape[2]
        nb = _f32(fuse_next[0], "next bias").reshape(-1)
        ns = _f32(fuse_next[1], "next scales").reshape(-1)
        nw = _f32(fuse_next[2], "next conv weight", (Cc, Cc))
        keep = keep + (nb, ns, nw)
        a.mix.next_actnorm_bias, a.mix.next_actnorm_scales, a.mix.next_conv_weight = _ptr(nb), _ptr(ns), _ptr(nw)
        if next_mask is not None:
            nm = _f32(next_mask, "next mask").reshape(-1)
            if nm.numel() != Cc:
                raise ValueError("next_mask has %d entries, expected %d" % (nm.numel(), Cc))
            z_masked = torch.empty_like(z)
            keep = keep + (nm,)
            a.next_mask, a.z_masked_out = _ptr(nm), _ptr(z_masked)
    elif next_mask is not None:
        raise ValueError("next_mask needs fuse_next")
    _call("cnf_linear_mixcdf_inv" if reverse else "cnf_linear_mixcdf_fwd", a, z, keep)
    if z_masked is not None:
        return z_out, ldj_t, reg, z_masked
    return z_out, ldj_t, reg


# ----------------------------------------------------------------------------------------------
# glue of the graph coupling networks
# ----------------------------------------------------------------------------------------------
def layernorm(x, weight, bias, eps=1e-5):
    """nn.LayerNorm over the last dimension (``cnf_layernorm``), one warp per row."""
    x = _f32(x, "x")
    H = x.shape[-1]
    weight, bias = _f32(weight, "weight", (H,)), _f32(bias, "bias", (H,))
    y = torch.empty_like(x)
    a = L.LayernormArgs()
    a.M, a.H = x.numel() // H, H
    a.x, a.gamma, a.beta, a.eps, a.y = _ptr(x), _ptr(weight), _ptr(bias), float(eps), _ptr(y)
    _call("cnf_layernorm", a, x, (x, weight, bias))
    return y


def _adjacency(adjacency, B, N):
    if not adjacency.is_cuda:
        raise RuntimeError("categoricalnf_b200: adjacency lives on %s - CUDA only" % adjacency.device)
    if adjacency.dtype != torch.int64:
        adjacency = adjacency.long()
    if tuple(adjacency.shape) != (B, N, N):
        raise ValueError("adjacency has shape %s, expected %s" % (tuple(adjacency.shape), (B, N, N)))
    return adjacency.contiguous()


def _rows(t, name, B, N):
    """[B,N,F] or [B*N,F] features, possibly a column slice of a wider row-major matrix -> (tensor, row pitch in floats)."""
    if not t.is_cuda:
        raise RuntimeError("categoricalnf_b200: %s lives on %s - CUDA only" % (name, t.device))
    if t.dtype != torch.float32:
        t = t.float()
    if t.dim() == 3:
        if t.stride(2) == 1 and t.stride(0) == N * t.stride(1):
            return t, t.stride(1)
        t = t.contiguous()
        return t, t.shape[-1]
    if t.dim() == 2 and t.stride(1) == 1:
        return t, t.stride(0)
    t = t.contiguous()
    return t, t.shape[-1]


def graph_attention_aggregate(hs, hr, attn_weight, adjacency, num_edges, *, leaky_slope=0.2, activation=None, scores=None):
    """RelationGraphAttention core (graph_layers.py:92-151): ``hs`` [B,N,H*Dh], ``hr`` [B,N,(E+1)*H*Dh] (either may be a
    column slice of one wider projection output), ``attn_weight`` [H,2,Dh], ``adjacency`` [B,N,N] integer edge types
    (0 = none).  Returns the attention output [B,N,H*Dh] (GELU applied when ``activation="gelu"``).
    ``scores = (score_s [B*N,H], score_r [B*N,(E+1)*H])`` (column slices allowed) skips ``cnf_graph_attn_scores`` - the
    logits are linear in the projection input, so the caller may have produced them as extra projection columns."""
    B, N = hs.shape[0], hs.shape[1]
    H, _, Dh = attn_weight.shape
    E = int(num_edges)
    adjacency = _adjacency(adjacency, B, N)
    hs_t, ld_hs = _rows(hs, "hs", B, N)
    hr_t, ld_hr = _rows(hr, "hr", B, N)
    if hs.shape[-1] != H * Dh or hr.shape[-1] != (E + 1) * H * Dh:
        raise ValueError("hs / hr have %d / %d features, expected %d / %d" % (hs.shape[-1], hr.shape[-1], H * Dh, (E + 1) * H * Dh))
    dev = hs.device
    if scores is None:
        aw = _f32(attn_weight, "attn_weight")
        score_s = torch.empty(B * N, H, dtype=torch.float32, device=dev)
        score_r = torch.empty(B * N, (E + 1) * H, dtype=torch.float32, device=dev)
        a = L.GraphAttnScoresArgs()
        a.M, a.E, a.H, a.Dh = B * N, E, H, Dh
        a.hs, a.hr, a.ld_hs, a.ld_hr = _ptr(hs_t), _ptr(hr_t), ld_hs, ld_hr
        a.attn_weight, a.score_s, a.score_r = _ptr(aw), _ptr(score_s), _ptr(score_r)
        _call("cnf_graph_attn_scores", a, hs_t, (hs_t, hr_t, aw))
        ld_ss, ld_sr = H, (E + 1) * H
    else:
        score_s, ld_ss = _rows(scores[0], "score_s", B, N)
        score_r, ld_sr = _rows(scores[1], "score_r", B, N)
        if scores[0].shape[-1] != H or scores[1].shape[-1] != (E + 1) * H:
            raise ValueError("scores have %d / %d columns, expected %d / %d" % (scores[0].shape[-1], scores[1].shape[-1], H, (E + 1) * H))
    out = torch.empty(B, N, H * Dh, dtype=torch.float32, device=dev)
    g = L.GraphAggregateArgs()
    g.B, g.N, g.E, g.H, g.Dh = B, N, E, H, Dh
    g.adjacency, g.hs, g.hr, g.ld_hs, g.ld_hr = _ptr(adjacency), None, _ptr(hr_t), ld_hs, ld_hr
    g.score_s, g.score_r, g.ld_score_s, g.ld_score_r, g.num_neighbours = _ptr(score_s), _ptr(score_r), ld_ss, ld_sr, None
    g.mode, g.leaky_slope, g.activation, g.out = 1, float(leaky_slope), ACTIVATION[activation], _ptr(out)
    _call("cnf_graph_aggregate", g, hs_t, (adjacency, hr_t, score_s, score_r))
    return out


def graph_mean_aggregate(hs, hr, adjacency, num_edges, num_neighbours=None, *, activation=None):
    """RelationGraphConv core (graph_layers.py:43-50): ``hs`` [B,N,C] + sum over neighbours j of ``hr`` [B,N,E*C] rows
    (block of the edge type) divided by ``num_neighbours`` [B,N] (None -> the number of edges found)."""
    B, N, Cc = hs.shape
    E = int(num_edges)
    adjacency = _adjacency(adjacency, B, N)
    hs_t, ld_hs = _rows(hs, "hs", B, N)
    hr_t, ld_hr = _rows(hr, "hr", B, N)
    if hr.shape[-1] != E * Cc:
        raise ValueError("hr has %d features, expected %d" % (hr.shape[-1], E * Cc))
    nn_t = _opt_f32(num_neighbours, "num_neighbours", (B, N))
    out = torch.empty(B, N, Cc, dtype=torch.float32, device=hs.device)
    g = L.GraphAggregateArgs()
    g.B, g.N, g.E, g.H, g.Dh = B, N, E, 1, Cc
    g.adjacency, g.hs, g.hr, g.ld_hs, g.ld_hr = _ptr(adjacency), _ptr(hs_t), _ptr(hr_t), ld_hs, ld_hr
    g.score_s, g.score_r, g.ld_score_s, g.ld_score_r, g.num_neighbours = None, None, 0, 0, _ptr(nn_t)
    g.mode, g.leaky_slope, g.activation, g.out = 0, 0.0, ACTIVATION[activation], _ptr(out)
    _call("cnf_graph_aggregate", g, hs_t, (adjacency, hs_t, hr_t, nn_t))
    return out


def skip_gate(orig, skip, config):
    """GNNSkipConnection combination (graph_layers.py:722-733); ``skip`` = skip_layer(feat)."""
    orig = _f32(orig, "orig")
    H = orig.shape[-1]
    skip = _f32(skip, "skip", orig.shape[:-1] + ((H if config == 0 else 2 * H),))
    out = torch.empty_like(orig)
    a = L.SkipGateArgs()
    a.M, a.H, a.config = orig.numel() // H, H, int(config)
    a.orig, a.skip, a.out = _ptr(orig), _ptr(skip), _ptr(out)
    _call("cnf_skip_gate", a, orig, (orig, skip))
    return out


def _rows2(t, name):
    """2-D feature rows, possibly a column slice of a wider row-major matrix -> (tensor, pitch)."""
    if not t.is_cuda:
        raise RuntimeError("categoricalnf_b200: %s lives on %s - CUDA only" % (name, t.device))
    if t.dtype != torch.float32:
        t = t.float()
    if t.dim() != 2:
        t = t.reshape(-1, t.shape[-1])
    if t.stride(1) != 1 or (t.shape[0] > 1 and t.stride(0) < t.shape[1]):
        t = t.contiguous()
    return t, (t.stride(0) if t.shape[0] > 1 else t.shape[1])


def edge_aggregate(rev, node_val, edge_val, edge_logit, num_heads, *, mode="sigmoid", node_q=None, node_k=None, scale=1.0):
    """Edge -> node attention of the Edge-GNN (``cnf_edge_aggregate``).  ``rev`` [B,P] int64 (1 + compact row of the pair,
    0 = invalid), ``node_val`` [B*N, H*Dh], ``edge_val`` [R, H*Dh], ``edge_logit`` [R, H] (column slices allowed);
    ``mode="qkv"`` additionally takes ``node_q`` / ``node_k`` [B*N, H*Dh].  Returns [B*N, H*Dh]."""
    B, P = rev.shape
    N = int(round((1 + (1 + 8 * P) ** 0.5) / 2))
    if N * (N - 1) // 2 != P:
        raise ValueError("rev has %d pair columns, not N(N-1)/2" % P)
    H = int(num_heads)
    HD = node_val.shape[-1]
    Dh = HD // H
    rev = rev.contiguous()
    nv, ld_nv = _rows2(node_val, "node_val")
    ev, ld_ev = _rows2(edge_val, "edge_val")
    el, ld_el = _rows2(edge_logit, "edge_logit")
    if nv.shape[0] != B * N or ev.shape[-1] != HD or el.shape[-1] != H:
        raise ValueError("edge_aggregate: inconsistent shapes")
    a = L.EdgeAggregateArgs()
    a.B, a.N, a.H, a.Dh, a.R = B, N, H, Dh, ev.shape[0]
    keep = [rev, nv, ev, el]
    a.rev, a.node_val, a.edge_val, a.edge_logit = _ptr(rev), _ptr(nv), _ptr(ev), _ptr(el)
    a.ld_node_val, a.ld_edge_val, a.ld_edge_logit = ld_nv, ld_ev, ld_el
    if mode == "qkv":
        q, ld_q = _rows2(node_q, "node_q")
        k, ld_k = _rows2(node_k, "node_k")
        keep += [q, k]
        a.node_q, a.node_k, a.ld_node_q, a.ld_node_k, a.mode = _ptr(q), _ptr(k), ld_q, ld_k, 1
    else:
        a.mode = 0
    a.scale = float(scale)
    out = torch.empty(B * N, HD, dtype=torch.float32, device=nv.device)
    a.out = _ptr(out)
    _call("cnf_edge_aggregate", a, nv, keep)
    return out


def pair_combine(flat_indices, x_indices, edge_lin, node_lin, num_nodes, *, activation="gelu"):
    """Node -> edge message of the Edge-GNN (``cnf_pair_combine``): ``act(edge_lin[r] + node_lin[b,x1[p]] + node_lin[b,x2[p]])``
    for every compact pair row r (``flat_indices[r] = b*P + p``)."""
    el, ld_e = _rows2(edge_lin, "edge_lin")
    nl, ld_n = _rows2(node_lin, "node_lin")
    x1, x2 = x_indices[0].contiguous(), x_indices[1].contiguous()
    flat_indices = flat_indices.contiguous()
    He = el.shape[-1]
    out = torch.empty(el.shape[0], He, dtype=torch.float32, device=el.device)
    a = L.PairCombineArgs()
    a.R, a.N, a.He = el.shape[0], int(num_nodes), He
    a.flat_indices, a.x_indices1, a.x_indices2 = _ptr(flat_indices), _ptr(x1), _ptr(x2)
    a.edge_lin, a.node_lin, a.ld_edge, a.ld_node = _ptr(el), _ptr(nl), ld_e, ld_n
    a.activation, a.out = ACTIVATION[activation], _ptr(out)
    if el.shape[0] > 0:
        _call("cnf_pair_combine", a, nl, (flat_indices, x1, x2, el, nl))
    return out
