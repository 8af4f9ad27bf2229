from .help_layers import LinearNet, SimpleLinearLayer, run_sequential_with_mask
from .linear import TCLinear, convert_linears, split_final_linear
from .graph_layers import GNNSkipConnection, RelationGraphAttention, RelationGraphConv, RGCNNet

__all__ = ["LinearNet", "SimpleLinearLayer", "run_sequential_with_mask", "TCLinear", "convert_linears",
           "split_final_linear", "GNNSkipConnection", "RelationGraphAttention", "RelationGraphConv", "RGCNNet"]
