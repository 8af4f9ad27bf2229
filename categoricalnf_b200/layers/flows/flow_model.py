"""Ordered container of flow layers with the running log-det-Jacobian accumulator
(reference layers/flows/flow_model.py:9-146).

Calling convention kept from the reference: every layer is invoked as
``layer(z, reverse=..., get_ldj_per_layer=..., **kwargs)`` *without* the running ldj (:34) and
returns its own contribution, which is added here (:44).  The reference synchronises with the
host once per layer for its NaN assert (:42); the kernels instead OR health bits into a device
status word that is read once per call (``ops.check_status``), raising the same AssertionError.
"""
import torch
import torch.nn as nn

from ... import ops
from .activation_normalization import ActNormFlow
from .permutation_layers import InvertibleConv


def _actnorm_conv_fused(z, an, conv, out_mask, channel_padding_mask=None, length=None, ldj_acc=None, **kwargs):
    """``ActNormFlow`` + ``InvertibleConv`` of one flow block in ONE pass over z (``cnf_invconv_apply`` with the
    ActNorm prologue), optionally emitting ``z_out * out_mask`` = the masked input of the coupling that follows.
    Returns (z_out, ldj of both layers, {}[, z_masked]).  ``ldj_acc`` given: the two layers' ldj - a per-sample constant
    times the sample's length - is added into that running accumulator by one ``cnf_ldj_axpy`` (with the constant cached
    per parameter version when there is no padding) and None is returned in its place."""
    from .mixture_cdf_layer import add_next_block_ldj
    weight, sldj = conv._get_weight(device_name=str(z.device), inverse=False)
    out = ops.invconv_apply(z, weight, sldj, None, pad=channel_padding_mask, pre_actnorm=(an.bias, an.scales),
                            out_mask=out_mask)
    if ldj_acc is not None and length is None and channel_padding_mask is None:
        key = (ops.param_epoch(), an.scales._version, an.scales.data_ptr(), sldj.data_ptr(), sldj._version)
        hit = an.__dict__.get("_cnf_block_const")
        if hit is None or hit[0] != key:
            hit = (key, (an.scales.detach().sum() + sldj.reshape(())).reshape(1))
            an.__dict__["_cnf_block_const"] = hit
        ops.ldj_axpy(ldj_acc, alpha=float(z.size(1)), alpha_dev=hit[1])
        ldj = None
    else:
        ldj = torch.zeros(z.size(0), dtype=torch.float32, device=z.device) if ldj_acc is None else ldj_acc
        add_next_block_ldj(ldj, an, sldj, z.size(1), channel_padding_mask, length)
        ldj = None if ldj_acc is not None else ldj
    return (out[0], ldj, {}) + ((out[2],) if out_mask is not None else ())


def _next_coupling_mask(order, pos):
    """Channel mask [1, C] of the layer at ``pos`` when that layer accepts a pre-masked input, else None."""
    nxt = order[pos][1] if pos < len(order) else None
    if getattr(nxt, "accepts_masked_input", False) and nxt.mask.dim() == 2 and nxt.mask.size(0) == 1:
        needs = getattr(nxt, "needs_masked_input", None)
        if needs is not None and not needs():
            return None           # the layer folds its mask into its projection weight: no masked copy of z is needed
        return nxt.mask
    return None


class FlowModel(nn.Module):

    def __init__(self, layers=None, name="Flow model"):
        super().__init__()
        self.flow_layers = nn.ModuleList()
        self.name = name
        self.fuse_blocks = True   # evaluation-time kernel fusion across layer boundaries (results identical)
        if layers is not None:
            self.add_layers(layers)

    def add_layers(self, layers):
        for layer in layers:
            self.flow_layers.append(layer)
        self.print_overview()

    def forward(self, z, ldj=None, reverse=False, get_ldj_per_layer=False, check_nan=True, **kwargs):
        own_ldj = ldj is None
        if ldj is None:
            ldj = z.new_zeros(z.size(0), dtype=torch.float32)
        order = list(enumerate(self.flow_layers))
        if reverse:
            order.reverse()
        ldj_per_layer = []
        skip = 0
        can_fuse = self.fuse_blocks and not reverse and not get_ldj_per_layer and not torch.is_grad_enabled() and z.is_cuda
        # evaluation-time fusion also lets kernels add their ldj straight into the running accumulator (never into a
        # caller-provided tensor: that one is copied once)
        acc_mode = can_fuse
        if acc_mode and not own_ldj:
            ldj = ldj.float().clone()
        masked = None   # z * mask of the upcoming coupling layer, when the previous kernel already produced it
        for pos, (index, layer) in enumerate(order):
            if skip > 0:
                skip -= 1
                continue
            res = None
            extra = {}
            if masked is not None and getattr(layer, "accepts_masked_input", False):
                extra["cnf_masked_input"] = masked
            masked = None
            if can_fuse and hasattr(layer, "try_forward_fused") and pos + 2 < len(order):
                # [encoding | mixture coupling] followed by ActNorm + 1x1 conv: the two bandwidth-only
                # layers run inside the producing kernel's epilogue (two passes over z saved per block)
                an, conv = order[pos + 1][1], order[pos + 2][1]
                if type(an) is ActNormFlow and type(conv) is InvertibleConv and not an.training and not conv.training:
                    nmask = _next_coupling_mask(order, pos + 3) if getattr(layer, "accepts_masked_input", False) else None
                    if nmask is not None:
                        extra["cnf_next_mask"] = nmask
                    res = layer.try_forward_fused(z, an, conv, **extra, **kwargs)
                    extra.pop("cnf_next_mask", None)
                    if res is not None:
                        skip = 2
            if res is None and can_fuse and type(layer) is ActNormFlow and pos + 1 < len(order) and z.dim() == 3 \
                    and type(order[pos + 1][1]) is InvertibleConv and not layer.training and not order[pos + 1][1].training:
                # ActNorm + 1x1 conv of a block in one pass, plus the masked input of the coupling that follows
                res = _actnorm_conv_fused(z, layer, order[pos + 1][1], _next_coupling_mask(order, pos + 2),
                                          ldj_acc=ldj if acc_mode else None, **kwargs)
                skip = 1
            if res is not None and len(res) == 4:
                masked = res[3]
                res = res[:3]
            if res is None and acc_mode and hasattr(layer, "forward_accumulate") and not layer.training:
                z_acc = layer.forward_accumulate(z, ldj, **extra, **kwargs)      # ldj added in the kernel
                if z_acc is not None:
                    z = z_acc
                    ldj_per_layer.append({})
                    continue
            if res is None:
                res = layer(z, reverse=reverse, get_ldj_per_layer=get_ldj_per_layer, **extra, **kwargs)
            if len(res) == 2:
                z, layer_ldj = res
                detail = layer_ldj
            elif len(res) == 3:
                z, layer_ldj, detail = res
            else:
                raise ValueError("[!] ERROR: Got more return values than expected: %i (layer %i)" % (len(res), index + 1))
            if layer_ldj is not None:       # None: the kernel(s) accumulated into `ldj` already
                ldj = ldj + layer_ldj
            if isinstance(detail, list):
                ldj_per_layer += detail
            else:
                ldj_per_layer.append(detail)
        if check_nan and z.is_cuda:
            # one lazy host read for the whole stack instead of one assert per layer (:42)
            ops.check_status(z.device, where=self.name)
        if get_ldj_per_layer:
            return z, ldj, ldj_per_layer
        return z, ldj

    def reverse(self, z):
        # upstream passes the bound method as ldj here (App. B #6); the intent is plain inversion
        return self.forward(z, reverse=True)

    def test_reversibility(self, z, **kwargs):
        """Forward then inverse of every layer; prints which layers do not reproduce their input
        exactly (:60-78).  Returns True when all layers reproduced bit-exactly."""
        failed = False
        for index, layer in enumerate(self.flow_layers):
            z_fwd, ldj_fwd = layer(z, reverse=False, **kwargs)[:2]
            z_rec, ldj_rec = layer(z_fwd, reverse=True, **kwargs)[:2]
            if (z_fwd - z_rec).abs().sum() != 0 or (ldj_fwd + ldj_rec).abs().sum() != 0:
                print("-" * 100)
                print("[!] WARNING: Reversibility check failed for layer index %i" % index)
                print(layer.info())
                print("-" * 100)
                failed = True
        print("+" * 100)
        print("Reversibility test %s (tested %i layers)" % ("failed" if failed else "succeeded", len(self.flow_layers)))
        print("+" * 100)
        return not failed

    def get_inner_activations(self, z, reverse=False, return_names=False, **kwargs):
        outs, names = [z.detach()], []
        for layer in (reversed(self.flow_layers) if reverse else self.flow_layers):
            z = layer(z, reverse=reverse, **kwargs)[0]
            outs.append(z.detach())
            names.append(layer.__class__.__name__)
        return (outs, names) if return_names else outs

    # -- data-dependent initialisation (:95-135) ----------------------------------------------------
    def initialize_data_dependent(self, batch_list):
        """``batch_list``: list of ``(z, kwargs)`` tuples, pushed through the stack layer by layer."""
        with torch.no_grad():
            for index, layer in enumerate(self.flow_layers):
                print("Processing layer %i..." % (index + 1), end="\r")
                batch_list = FlowModel.run_data_init_layer(batch_list, layer)

    @staticmethod
    def run_data_init_layer(batch_list, layer):
        multi_input = isinstance(batch_list[0][0], (tuple, list))
        if layer.need_data_init():
            merged = {}
            for key in batch_list[0][1].keys():
                vals = [b[1][key] for b in batch_list]
                merged[key] = torch.cat(vals, dim=0) if isinstance(vals[0], torch.Tensor) else vals[0]
            if not multi_input:
                layer.data_init_forward(torch.cat([z for z, _ in batch_list], dim=0), **merged)
            else:
                n_in = len(batch_list[0][0])
                layer.data_init_forward(*[torch.cat([z[i] for z, _ in batch_list], dim=0) for i in range(n_in)],
                                        **merged)
        outs = []
        for z, kwargs in batch_list:
            if isinstance(z, (tuple, list)):
                res = layer(*z, reverse=False, **kwargs)
                cur = [e.detach() for e in res[:-1] if isinstance(e, torch.Tensor)]
                if len(res) == 4 and isinstance(res[-1], dict):
                    kwargs.update(res[-1])
                    cur = cur[:-1]
                outs.append(cur)
            else:
                outs.append(layer(z, reverse=False, **kwargs)[0].detach())
        return [(outs[i], batch_list[i][1]) for i in range(len(batch_list))]

    def need_data_init(self):
        return any(flow.need_data_init() for flow in self.flow_layers)

    def print_overview(self):
        lines = ["(%2i) %s" % (i + 1, layer.info()) for i, layer in enumerate(self.flow_layers)]
        width = max([20] + [len(s) for s in "\n".join(lines).split("\n")])
        print("=" * width)
        print("%s with %i flows" % (self.name, len(self.flow_layers)))
        print("-" * width)
        print("\n".join(lines))
        print("=" * width)
