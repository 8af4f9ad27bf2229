"""categoricalnf_b200 - B200-native (sm_100a) implementation of the coupling-layer hot path of
phlippe/CategoricalNF behind the reference's own ``FlowLayer`` API.

``categoricalnf_b200.ops``      tensor-level calls into the C-ABI CUDA library (include/cnf_b200.h)
``categoricalnf_b200.layers``   drop-in modules mirroring ``layers/flows`` and
                                ``layers/categorical_encoding`` of the reference
``categoricalnf_b200.install``  patch the modules into an importable reference checkout
``categoricalnf_b200.sharding`` one-process-per-GPU batch sharding + log-likelihood all-reduce
"""
__version__ = "0.1.0"
