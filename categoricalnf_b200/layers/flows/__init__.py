from .flow_layer import FlowLayer
from .flow_model import FlowModel
from .graphed import GraphedFlowForward
from .coupling_layer import CouplingLayer
from .mixture_cdf_layer import MixtureCDFCoupling
from .autoregressive_coupling import AutoregressiveMixtureCDFCoupling
from .activation_normalization import ActNormFlow, ExtActNormFlow
from .permutation_layers import InvertibleConv
from .node_edge_coupling import NodeEdgeCoupling, NodeEdgeFlowWrapper
from .sigmoid_layer import SigmoidFlow
from .distributions import LogisticDistribution, PriorDistribution, create_prior_distribution

__all__ = ["FlowLayer", "FlowModel", "GraphedFlowForward", "CouplingLayer", "MixtureCDFCoupling", "AutoregressiveMixtureCDFCoupling",
           "ActNormFlow", "ExtActNormFlow", "SigmoidFlow", "InvertibleConv", "NodeEdgeCoupling", "NodeEdgeFlowWrapper", "LogisticDistribution", "PriorDistribution",
           "create_prior_distribution"]
