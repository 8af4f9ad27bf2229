"""Backward passes of the fused layers.

Placeholder module: until the backward kernels are compiled into libcnf_b200.so every function
fails loudly.  There is deliberately no eager-PyTorch fallback on this path.
"""


def _missing(name):
    raise NotImplementedError(
        "categoricalnf_b200: the backward kernel of %s is not built yet; run this layer under "
        "torch.no_grad() (evaluation / sampling)" % name)


def mixcdf_backward(*a, **k):
    _missing("mixcdf")


def affine_backward(*a, **k):
    _missing("affine_coupling")


def actnorm_backward(*a, **k):
    _missing("actnorm")


def ext_actnorm_backward(*a, **k):
    _missing("ext_actnorm")


def invconv_backward(*a, **k):
    _missing("invconv")


def logistic_logprob_backward(*a, **k):
    _missing("logistic_logprob")
