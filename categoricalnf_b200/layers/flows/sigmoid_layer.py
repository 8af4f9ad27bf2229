"""Sigmoid flow (reference layers/flows/sigmoid_layer.py:12-51).

``z -> sigmoid(z)`` with ``ldj += sum(-z - 2 softplus(-z))``; reversed it is the logit of the input squeezed by
``alpha = 1e-5`` away from {0, 1}.  ``SigmoidFlow(reverse=True)`` swaps the two directions (``reverse_layer XOR
reverse``, :29).  One launch of ``cnf_sigmoid_flow`` (csrc/sigmoid_flow.cu) instead of ~8 eager ops and two NaN-assert
host syncs; numerical health goes to the device status word (``ops.check_status``).
"""
from ... import functional as CF
from .flow_layer import FlowLayer

ALPHA = 1e-5


class SigmoidFlow(FlowLayer):

    def __init__(self, reverse=False):
        super().__init__()
        self.reverse_layer = reverse

    def forward(self, z, ldj=None, reverse=False, sum_ldj=True, **kwargs):
        effective = (self.reverse_layer != reverse)
        return CF.sigmoid_flow(z, ldj, reverse=effective, alpha=ALPHA, sum_ldj=sum_ldj)

    def info(self):
        return "Sigmoid Flow"
