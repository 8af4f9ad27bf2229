"""Base class of every flow layer - the drop-in boundary (reference layers/flows/flow_layer.py:5-32).

Contract kept from the reference: ``forward(z, ldj=None, reverse=False, **kwargs)`` returns
``(z, ldj)`` or ``(z, ldj, detail)``; ``reverse`` is forward with ``reverse=True``;
``need_data_init`` / ``data_init_forward`` drive the data-dependent initialisation; ``info`` is a
one-line description used by ``FlowModel.print_overview``.
"""
import torch.nn as nn


class FlowLayer(nn.Module):

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
