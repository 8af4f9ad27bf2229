"""Logistic-mixture-CDF coupling (Flow++ style; reference layers/flows/mixture_cdf_layer.py).

``forward`` keeps the reference's signature, parameters (``scaling_factor``,
``mixture_scaling_factor``, ``nn.*``), its 3-tuple return and its quirk of *discarding* the
incoming ldj (App. B #1); the parameter split + transform + ldj reduction run as one
``cnf_mixcdf_fwd`` / ``cnf_mixcdf_inv`` launch (csrc/mixcdf.cu).
"""
import os

import torch
import torch.nn as nn

from ... import functional as CF
from ... import ops
from ..networks.linear import split_final_linear
from ._masks import mask_lists
from .coupling_layer import CouplingLayer


class MixtParams:
    """Opaque handle returned five times by :meth:`MixtureCDFCoupling.get_mixt_params`.

    The reference materialises five float64 tensors here (mixture_cdf_layer.py:145-180) and hands
    them straight to ``run_with_params`` (e.g. graph_node_edge_coupling.py:117-135,
    autoregressive_coupling.py:32-38).  The fused kernel consumes the raw network output, so the
    handle only records it; ``run_with_params`` recognises the handle and launches the kernel.
    """

    def __init__(self, nn_out, mask, num_mixtures, scaling_factor, mixture_scaling_factor):
        self.nn_out, self.mask, self.num_mixtures = nn_out, mask, num_mixtures
        self.scaling_factor, self.mixture_scaling_factor = scaling_factor, mixture_scaling_factor

    def materialize(self):
        """The five bounded, masked parameter tensors in float32 (diagnostics only)."""
        K, pn = self.num_mixtures, 2 + 3 * self.num_mixtures
        rec = self.nn_out.reshape(self.nn_out.shape[:-1] + (self.nn_out.shape[-1] // pn, pn))
        t, log_s = rec[..., 0], rec[..., 1]
        log_pi, mu, mls = rec[..., 2:2 + K], rec[..., 2 + K:2 + 2 * K], rec[..., 2 + 2 * K:]
        if self.scaling_factor is not None:
            b = self.scaling_factor.exp()
            log_s = torch.tanh(log_s / b.clamp(min=1.0)) * b
        if self.mixture_scaling_factor is not None:
            b = self.mixture_scaling_factor.exp()
            mls = torch.tanh(mls / b.clamp(min=1.0)) * b
        if self.mask is not None:
            keep = 1 - self.mask
            t, log_s = t * keep, log_s * keep
            log_pi, mu, mls = (v * keep.unsqueeze(-1) for v in (log_pi, mu, mls))
        return t, log_s, log_pi, mu, mls


def _mask_from_tensor(mask, z):
    """Broadcastable mask tensor ([1,1,C] / [1,S,1] / [1,C]) -> (mask_c, mask_s) host lists."""
    if mask is None:
        return None, None
    m = mask.detach().to("cpu", torch.float32)
    while m.dim() > 2 and m.shape[0] == 1:
        m = m[0]
    if m.dim() == 1:
        return m.tolist(), None
    if m.shape[0] == 1:
        return m.flatten().tolist(), None
    if m.shape[-1] == 1:
        return None, m.flatten().tolist()
    raise NotImplementedError("joint position x channel mask of shape %s" % (tuple(mask.shape),))


def add_next_block_ldj(ldj, actnorm, sldj, S, channel_padding_mask, length):
    """ldj[b] += sum(scales) * len_act[b] + sldj * len_conv[b]: the per-sample-constant terms of a fused
    ActNorm (activation_normalization.py:27-40: length, else unpadded positions, else S) and
    InvertibleConv (permutation_layers.py:111-117: length, else S)."""
    ssum = actnorm.scales.detach().sum().reshape(1)
    if length is not None:
        ops.ldj_axpy(ldj, alpha_dev=ssum + sldj.reshape(1), length=length.float())
    elif channel_padding_mask is None:
        ops.ldj_axpy(ldj, alpha=float(S), alpha_dev=ssum + sldj.reshape(1))
    else:
        ops.ldj_axpy(ldj, alpha_dev=ssum, length=channel_padding_mask.reshape(ldj.shape[0], -1).sum(dim=1))
        ops.ldj_axpy(ldj, alpha=float(S), alpha_dev=sldj.reshape(1))


class MixtureCDFCoupling(CouplingLayer):

    def __init__(self, c_in, mask, model_func, block_type=None, num_mixtures=10, regularizer_max=-1,
                 regularizer_factor=1, **kwargs):
        super().__init__(c_in=c_in, mask=mask, model_func=model_func, block_type=block_type,
                         c_out=c_in * (2 + num_mixtures * 3), **kwargs)
        self.num_mixtures = num_mixtures
        self.mixture_scaling_factor = nn.Parameter(torch.zeros(self.c_in, self.num_mixtures))
        self.regularizer_max = regularizer_max
        self.regularizer_factor = regularizer_factor

    # Evaluation-time fusion of the network's final projection with the transform (cnf_linear_mixcdf_*):
    # the [B,S,C*(2+3K)] network output is then never written to memory.  Results are identical to the
    # two-step path within the 3xTF32 projection error (~1e-6); set False to force the two-step path.
    fuse_final_projection = True
    projection_precision = "3xtf32"
    # The fused projection kernel can also run the next block's ActNorm + 1x1 conv in its epilogue, but the extra
    # shared memory costs it a pipeline stage (measured slower than ActNorm + conv as one separate pass), so the
    # container uses cnf_invconv_apply with the ActNorm prologue instead unless this is set.
    fuse_next_in_projection_kernel = os.environ.get("CNF_B200_FUSE_NEXT_IN_PROJECTION", "0") not in ("", "0")

    def _projection_split(self, z):
        """(features_fn, linear) when the final projection of ``self.nn`` can be fused for ``z``, else None."""
        if not (self.fuse_final_projection and z.is_cuda and z.dim() == 3) or torch.is_grad_enabled():
            return None
        split = split_final_linear(self.nn)
        if split is None:
            return None
        lin = split[1]
        if lin.out_features != self.c_in * (2 + 3 * self.num_mixtures) or lin.in_features % 4 != 0:
            return None
        return split

    accepts_masked_input = True   # FlowModel may hand over `cnf_masked_input` = z * mask made by the previous kernel

    def needs_masked_input(self):
        """False when the network IS its final Linear applied to the masked input (``nn.cnf_features_are_input``, e.g. a
        per-position linear conditioner): ``(z * mask) @ W^T == z @ (W * mask)^T`` exactly, so at evaluation time the mask
        is folded into the weight columns once per parameter version and no masked copy of z is ever made."""
        return not (getattr(self.nn, "cnf_features_are_input", False) and self.mask.dim() == 2 and self.mask.size(0) == 1
                    and not torch.is_grad_enabled() and self.fuse_final_projection)

    def _fused_shape_possible(self, z, lin, mask_c):
        """Static part of ``cnf_linear_mixcdf_fusable`` (csrc/linear_mixcdf.cu ``fused_shape_ok`` / ``check_fusable``):
        lets callers that cannot fall back cheaply decide before the network body has been evaluated."""
        C, K = z.size(2), self.num_mixtures
        tch = list(range(C)) if mask_c is None else [c for c, m in enumerate(mask_c) if float(m) == 0.0]
        if not tch or tch != list(range(tch[0], tch[0] + len(tch))):
            return False
        return (K, len(tch)) in ((8, 8), (8, 4), (16, 4), (4, 8), (4, 4)) and C % 4 == 0 and C <= 32 and lin.in_features % 4 == 0

    def _mask_folded_weight(self, lin):
        key = (ops.param_epoch(), lin.weight._version, lin.weight.data_ptr(), self.mask._version, self.mask.data_ptr())
        hit = self.__dict__.get("_cnf_folded")
        if hit is None or hit[0] != key:
            with torch.no_grad():
                w = (lin.weight * self.mask.reshape(1, -1).to(lin.weight.dtype)).contiguous()
            hit = (key, w)
            self.__dict__["_cnf_folded"] = hit
        return hit[1]

    def forward(self, z, ldj=None, reverse=False, channel_padding_mask=None, cnf_masked_input=None, **kwargs):
        # the incoming ldj is ignored and only this layer's ldj is returned, as upstream (:46-47,:63)
        res = self._forward_impl(z, reverse, channel_padding_mask, cnf_masked_input, None, kwargs)
        return res[:3]

    def forward_accumulate(self, z, ldj_acc, channel_padding_mask=None, cnf_masked_input=None, **kwargs):
        """Evaluation-time variant for ``FlowModel``: this layer's ldj is ADDED into ``ldj_acc`` [B] by the kernel itself
        (no per-layer ldj tensor, no separate add).  -> z_out, or None when the fused projection path is not available
        (the caller then uses :meth:`forward`)."""
        res = self._forward_impl(z, False, channel_padding_mask, cnf_masked_input, None, kwargs, ldj_acc=ldj_acc)
        return None if res is None else res[0]

    def _forward_impl(self, z, reverse, channel_padding_mask, masked_input, fuse, kwargs, ldj_acc=None):
        """``fuse`` = None or (actnorm, conv, next_mask): next-block epilogue of the fused projection kernel.
        ``ldj_acc``: accumulate the layer's ldj into this tensor (fused projection path only, else None is returned).
        Returns (z_out, ldj, detail[, z_masked]) - or None when ``fuse`` / ``ldj_acc`` was requested but is not possible."""
        mask_c, mask_s = mask_lists(self, "mask", z.size(1))
        split = self._projection_split(z)
        if (fuse is not None or ldj_acc is not None) and (split is None or not self._fused_shape_possible(z, split[1], mask_c)):
            return None           # decided BEFORE the network body runs: the caller falls back to forward()
        folded = split is not None and not self.needs_masked_input()
        if folded:
            x_in = z                      # mask folded into the weight columns below
        elif split is None and fuse is None and ldj_acc is None and self._train_fold_possible(z, mask_s):
            # training step with a per-position linear network: the mask goes into the (tiny) weight instead of two
            # elementwise passes over z (z * mask forward, grad * mask backward) - see _compact_projection
            compact = self._compact_projection(z, None, mask_c, mask_s, channel_padding_mask, kwargs) if not reverse else None
            if compact is not None:
                z_out, ldj, reg = CF.proj_mixcdf(z, compact[0], compact[1], compact[2], self.num_mixtures, self.scaling_factor,
                                                 self.mixture_scaling_factor, mask_c=mask_c, pad=channel_padding_mask,
                                                 reg_max=self.regularizer_max, reg_factor=self.regularizer_factor,
                                                 training=self.training, precision=self.projection_precision, linear_on_z=True)
                return z_out, ldj, {"ldj": ldj, "regularizer_ldj": reg}
            x_in = masked_input if masked_input is not None else z * self._prepare_mask(self.mask, z)
        else:
            x_in = masked_input if masked_input is not None else z * self._prepare_mask(self.mask, z)
        if split is not None:
            features_fn, lin = split
            feats = features_fn(x_in, **kwargs)
            weight = self._mask_folded_weight(lin) if folded else lin.weight
            if feats.dim() == 3 and ops.linear_mixcdf_fusable(z, feats, weight, self.num_mixtures, mask_c=mask_c, mask_s=mask_s):
                extra = {}
                if fuse is not None:
                    actnorm, conv, next_mask = fuse
                    weight_c, sldj = conv._get_weight(device_name=str(z.device), inverse=False)
                    extra = dict(fuse_next=(actnorm.bias, actnorm.scales, weight_c), next_mask=next_mask)
                out = ops.linear_mixcdf(
                    z, feats, weight, lin.bias, self.num_mixtures, mask_c=mask_c, mask_s=mask_s, pad=channel_padding_mask,
                    scaling_factor=self.scaling_factor, mixture_scaling_factor=self.mixture_scaling_factor, reverse=reverse,
                    reg_max=self.regularizer_max, reg_factor=self.regularizer_factor, training=self.training,
                    want_reg=ldj_acc is None, precision=self.projection_precision, ldj=ldj_acc, **extra)
                z_out, ldj, reg = out[:3]
                if ldj_acc is not None:
                    return (z_out, None, {})
                detail = {"ldj": ldj, "regularizer_ldj": reg}
                if fuse is not None:
                    detail = {"ldj": ldj.clone(), "regularizer_ldj": reg}
                    add_next_block_ldj(ldj, actnorm, sldj, z.size(1), channel_padding_mask, kwargs.get("length", None))
                return (z_out, ldj, detail) + tuple(out[3:])
            if fuse is not None or ldj_acc is not None:
                return None
            if folded:
                feats = features_fn(z * self._prepare_mask(self.mask, z), **kwargs)
            nn_out = lin(feats)     # shape / alignment outside the fused kernel: finish the network as usual
        else:
            if fuse is not None or ldj_acc is not None:
                return None
            compact = self._compact_projection(z, x_in, mask_c, mask_s, channel_padding_mask, kwargs) if not reverse else None
            if compact is not None:
                z_out, ldj, reg = CF.proj_mixcdf(z, compact[0], compact[1], compact[2], self.num_mixtures, self.scaling_factor,
                                                 self.mixture_scaling_factor, mask_c=mask_c, pad=channel_padding_mask,
                                                 reg_max=self.regularizer_max, reg_factor=self.regularizer_factor,
                                                 training=self.training, precision=self.projection_precision)
                return z_out, ldj, {"ldj": ldj, "regularizer_ldj": reg}
            nn_out = self.run_network(x=x_in, **kwargs)
        z_out, ldj, reg = CF.mixcdf(z, nn_out, self.num_mixtures, self.scaling_factor, self.mixture_scaling_factor,
                                    mask_c=mask_c, mask_s=mask_s, pad=channel_padding_mask, reverse=reverse,
                                    reg_max=self.regularizer_max, reg_factor=self.regularizer_factor,
                                    training=self.training)
        return z_out, ldj, {"ldj": ldj, "regularizer_ldj": reg}

    # Training-time counterpart of the fused projection: when the network ends in a Linear, only the weight rows of the
    # TRANSFORMED channels' records are multiplied - the transform never reads the conditioner half of the network output
    # (mixture_cdf_layer.py:166-173 zeroes it) - and the transform / its backward kernel take that compact
    # [B,S,Ct*(2+3K)] layout: half the projection GEMMs (forward, grad_x, grad_W) and no zeros written for conditioner
    # channels in dL/dnn_out.  Gradients of the skipped weight rows are exactly zero in the reference too.
    compact_projection_in_training = True

    def _train_fold_possible(self, z, mask_s):
        return (self.compact_projection_in_training and torch.is_grad_enabled() and z.is_cuda and z.dim() == 3 and mask_s is None
                and getattr(self.nn, "cnf_features_are_input", False) and self.mask.dim() == 2 and self.mask.size(0) == 1)

    def _compact_projection(self, z, x_in, mask_c, mask_s, channel_padding_mask, kwargs):
        """Operands of the compact projection - (features [B*S,H], weight rows and bias entries of the transformed channels'
        records) - or None when this configuration does not allow it.
        ``x_in`` None: the network is its final Linear on the masked input (``_train_fold_possible``) - the projection runs
        on the UNMASKED z with the mask folded into the weight columns, ``(z * m) W^T = z (W * m)^T``: the weight gradient
        flows back through that [rows, C] multiply, and the gradient wrt z comes out of the GEMM already masked."""
        if not (self.compact_projection_in_training and torch.is_grad_enabled() and z.is_cuda and z.dim() == 3):
            return None
        if mask_c is None or mask_s is not None:
            return None
        split = split_final_linear(self.nn)
        if split is None:
            return None
        features_fn, lin = split
        pn = 2 + 3 * self.num_mixtures
        if lin.out_features != self.c_in * pn or lin.in_features % 4 != 0:
            return None
        tch = [c for c, m in enumerate(mask_c) if float(m) == 0.0]
        if not tch or tch != list(range(tch[0], tch[0] + len(tch))) or len(tch) == self.c_in:
            return None                                   # needs a contiguous run, and a conditioner half worth skipping
        r0, r1 = tch[0] * pn, (tch[-1] + 1) * pn
        if (r0 * lin.in_features) % 4 != 0 or ((r1 - r0) % 4) != 0:
            return None                                   # 16-byte aligned weight block / output rows
        probe = z.new_empty(z.shape[0], z.shape[1], r1 - r0)
        if ops.mixcdf_path(z, probe, self.num_mixtures, mask_c=mask_c, compact=True) == "generic":
            return None
        weight = lin.weight[r0:r1]
        if x_in is None:
            if lin.in_features != z.shape[-1]:
                return None
            feats = z
            weight = weight * self.mask.reshape(1, -1).to(weight.dtype)
        else:
            feats = features_fn(x_in, **kwargs)
        if feats.dim() != 3:
            return None
        bias = None if lin.bias is None else lin.bias[r0:r1]
        weight._cnf_cache_lo = True       # a weight block: its 3xTF32 split is made once per call, not once per tile (ops.weight_split)
        # (no pad multiply on the network output: channel_padding_mask is bound by forward() and never reaches run_network,
        # App. B #3).  The projection and the transform run as one autograd node (CF.proj_mixcdf).
        return feats.reshape(-1, feats.shape[-1]), weight, bias

    def try_forward_fused(self, z, actnorm, conv, channel_padding_mask=None, length=None, cnf_masked_input=None,
                          cnf_next_mask=None, **kwargs):
        """Forward of this layer AND of the following ``ActNormFlow`` + ``InvertibleConv`` in one kernel
        (evaluation only).  Returns ``(z, ldj, detail[, z_masked])`` with the three layers' ldj summed
        (``z_masked`` = z * ``cnf_next_mask`` when the kernel produced it), or None when the kernel cannot fuse
        for this shape / mask (caller then runs the layers one by one)."""
        if self.training or z.dim() != 3:
            return None
        if self._projection_split(z) is not None:
            if not self.fuse_next_in_projection_kernel:
                return None
            # final projection + transform + next block in the tcgen05 kernel
            if length is not None:
                kwargs = dict(kwargs, length=length)
            return self._forward_impl(z, False, channel_padding_mask, cnf_masked_input, (actnorm, conv, cnf_next_mask), kwargs)
        mask_c, mask_s = mask_lists(self, "mask", z.size(1))
        # cheap shape pre-check (the kernel fuses for C=16, K=8, 8 contiguous transformed channels) so the
        # network is not run twice; the library has the final word below
        if z.size(2) != 16 or self.num_mixtures != 8 or mask_s is not None or mask_c is None or \
                sum(1 for m in mask_c if m == 0) != 8:
            return None
        x_in = cnf_masked_input if cnf_masked_input is not None else z * self._prepare_mask(self.mask, z)
        nn_out = self.run_network(x=x_in, length=length, **kwargs)
        if not ops.mixcdf_fusable(z, nn_out, self.num_mixtures, mask_c=mask_c, mask_s=mask_s):
            return None   # note: the network has run; the caller's unfused call runs it again
        weight, sldj = conv._get_weight(device_name=str(z.device), inverse=False)
        z_out, ldj, reg = ops.mixcdf(z, nn_out, self.num_mixtures, mask_c=mask_c, mask_s=mask_s, pad=channel_padding_mask,
                                     scaling_factor=self.scaling_factor, mixture_scaling_factor=self.mixture_scaling_factor,
                                     reg_max=self.regularizer_max, reg_factor=self.regularizer_factor, training=False,
                                     want_reg=True, fuse_next=(actnorm.bias, actnorm.scales, weight))
        detail = {"ldj": ldj.clone(), "regularizer_ldj": reg}
        add_next_block_ldj(ldj, actnorm, sldj, z.size(1), channel_padding_mask, length)
        return z_out, ldj, detail

    # -- static entry points used by other files ----------------------------------------------------
    @staticmethod
    def get_mixt_params(nn_out, mask, num_mixtures, scaling_factor=None, mixture_scaling_factor=None):
        h = MixtParams(nn_out, mask, num_mixtures, scaling_factor, mixture_scaling_factor)
        return h, h, h, h, h

    @staticmethod
    def run_with_params(orig_z, t, log_s, log_pi, mixt_t, mixt_log_s, reverse=False, reg_max=-1, reg_factor=1,
                        mask=None, channel_padding_mask=None, is_training=True, return_reg_ldj=False):
        """Fused transform for parameters obtained from :meth:`get_mixt_params` (handle) or given
        as explicit tensors (packed into the record layout, bounding disabled)."""
        z = orig_z.float()
        if isinstance(t, MixtParams):
            h = t
            nn_out, K, sf, msf, pre = h.nn_out, h.num_mixtures, h.scaling_factor, h.mixture_scaling_factor, False
            if mask is None:
                mask = h.mask
        else:
            K = log_pi.shape[-1]
            nn_out = torch.cat([t.unsqueeze(-1), log_s.unsqueeze(-1), log_pi, mixt_t, mixt_log_s], dim=-1).float()
            nn_out = nn_out.reshape(z.shape[:-1] + (z.shape[-1] * (2 + 3 * K),))
            sf = msf = None
            pre = True
        mask_c, mask_s = _mask_from_tensor(mask, z)
        z_out, ldj, reg = CF.mixcdf(z, nn_out, K, sf, msf, mask_c=mask_c, mask_s=mask_s, pad=channel_padding_mask,
                                    reverse=reverse, reg_max=reg_max, reg_factor=reg_factor, training=is_training,
                                    prebounded=pre)
        # upstream multiplies by the padding mask only in the callers (:76); the kernel has already
        # done so, which is idempotent for 0/1 masks.
        if return_reg_ldj:
            # upstream returns the regulariser per element [B,S,C] and its callers reduce over dims 1.. (e.g.
            # graph_node_edge_coupling.py:138-139); the kernel has reduced already, so hand back [B,1,1] - the same
            # reduction is then a no-op (a bare [B] would make `sum(dim=[])` collapse the batch as well).
            return z_out, ldj, (reg.reshape((-1,) + (1,) * (z.dim() - 1)) if not reverse else None)
        return z_out, ldj

    def info(self):
        kind = "channel" if self.mask.size(0) == 1 else "chess"
        text = "Mixture CDF Coupling Layer - Input size %i" % self.c_in
        if self.block_type is not None:
            text += ", block type %s" % self.block_type
        return text + ", %i mixtures, mask ratio %.2f, %s mask" % (self.num_mixtures, (1 - self.mask).mean().item(), kind)
