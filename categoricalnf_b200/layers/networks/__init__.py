from .help_layers import LinearNet, SimpleLinearLayer, run_sequential_with_mask

__all__ = ["LinearNet", "SimpleLinearLayer", "run_sequential_with_mask"]
