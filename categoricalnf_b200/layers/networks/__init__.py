from .help_layers import LinearNet, SimpleLinearLayer, run_sequential_with_mask
from .linear import TCLinear, convert_linears, split_final_linear
from .graph_layers import (Edge2NodeAttnLayer, Edge2NodeQKVAttnLayer, EdgeGNN, EdgeGNNLayer, GNNSkipConnection, Node2EdgePlainLayer,
                           RelationGraphAttention, RelationGraphConv, RGCNNet)

__all__ = ["LinearNet", "SimpleLinearLayer", "run_sequential_with_mask", "TCLinear", "convert_linears",
           "split_final_linear", "GNNSkipConnection", "RelationGraphAttention", "RelationGraphConv", "RGCNNet",
           "Edge2NodeAttnLayer", "Edge2NodeQKVAttnLayer", "EdgeGNN", "EdgeGNNLayer", "Node2EdgePlainLayer"]
