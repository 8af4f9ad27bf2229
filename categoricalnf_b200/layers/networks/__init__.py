from .help_layers import LinearNet, SimpleLinearLayer, run_sequential_with_mask
from .linear import TCLinear, convert_linears, split_final_linear

__all__ = ["LinearNet", "SimpleLinearLayer", "run_sequential_with_mask", "TCLinear", "convert_linears",
           "split_final_linear"]
