"""Host-side view of coupling masks.

The kernels take the mask as two small host arrays (conditioner channels, conditioner positions
of a periodic chess pattern - include/cnf_b200.h ``cnf_mask``).  The module keeps the mask as a
registered buffer exactly like the reference (coupling_layer.py:20), so this helper derives the
host arrays once per buffer version and caches them on the module.
"""
import math

import torch


def mask_lists(module, attr="mask", seq_len=None):
    """-> (mask_c list | None, mask_s list | None) for buffer ``module.<attr>``.

    ``[1, C]`` is a channel mask; ``[S_m, 1]`` a chess mask over positions that the reference tiles
    along the sequence (coupling_layer.py:67-74), i.e. position s uses entry ``s % S_m``.
    """
    mask = getattr(module, attr)
    key = (attr, mask.data_ptr(), mask._version, tuple(mask.shape))
    cache = module.__dict__.setdefault("_cnf_mask_cache", {})
    hit = cache.get(attr)
    if hit is None or hit[0] != key:
        host = mask.detach().to("cpu", torch.float32)
        if host.dim() != 2:
            raise ValueError("coupling mask must be 2-D ([1,C] or [S,1]), got %s" % (tuple(host.shape),))
        if host.shape[0] == 1:
            val = (host.flatten().tolist(), None)
        elif host.shape[1] == 1:
            val = (None, host.flatten().tolist())
        else:
            raise NotImplementedError("joint position x channel masks %s are not supported" % (tuple(host.shape),))
        hit = (key, val)
        cache[attr] = hit
    mask_c, mask_s = hit[1]
    if mask_s is not None and seq_len is not None and len(mask_s) > seq_len:
        mask_s = mask_s[:seq_len]      # reference truncates a longer mask (coupling_layer.py:72-73)
    return mask_c, mask_s


def broadcast_mask(mask, z):
    """The tensor form used to blank the conditioner input ``z * mask`` (coupling_layer.py:67-74)."""
    m = mask.unsqueeze(0) if z.dim() > mask.dim() else mask
    S = z.size(1)
    if 1 < m.size(1) < S:
        m = m.repeat(1, int(math.ceil(S / m.size(1))), 1)
    if m.size(1) > S:
        m = m[:, :S]
    return m
