"""Callers of the hot path that are rebuilt on the drop-in layers: the flow models of the
reference's ``experiments/`` whose coupling networks run on the sm_100a kernels.  Training loops, datasets and
CLIs stay the reference's own."""
