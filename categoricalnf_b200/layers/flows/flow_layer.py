"""Base class of every flow layer - the drop-in boundary (reference layers/flows/flow_layer.py:5-32).

Contract kept from the reference: ``forward(z, ldj=None, reverse=False, **kwargs)`` returns
``(z, ldj)`` or ``(z, ldj, detail)``; ``reverse`` is forward with ``reverse=True``;
``need_data_init`` / ``data_init_forward`` drive the data-dependent initialisation; ``info`` is a
one-line description used by ``FlowModel.print_overview``.
"""
import torch.nn as nn

from ... import ops


class FlowLayer(nn.Module):

    def train(self, mode=True):
        # caches derived from parameters (built 1x1-conv matrices, fused / split projection weights, captured CUDA
        # graphs) are keyed on ops.param_epoch(): a train <-> eval switch starts a new generation
        if bool(mode) != self.training:
            ops.invalidate_caches()
        return super().train(mode)

    def forward(self, z, ldj=None, reverse=False, **kwargs):
        raise NotImplementedError

    def reverse(self, z, ldj=None, **kwargs):
        return self.forward(z, ldj, reverse=True, **kwargs)

    def need_data_init(self):
        return False

    def data_init_forward(self, input_data, **kwargs):
        raise NotImplementedError

    def info(self):
        raise NotImplementedError
