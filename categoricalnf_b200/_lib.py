"""ctypes binding of the C ABI declared in ``include/cnf_b200.h``.

The shared library is built in-tree by :mod:`categoricalnf_b200.build` (``libcnf_b200.so`` next to
this file).  There is deliberately no fallback: if the library is missing or a symbol does not
resolve, importing the compute ops raises.
"""
from __future__ import annotations

import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libcnf_b200.so")

c_f32p = C.POINTER(C.c_float)
c_i64p = C.POINTER(C.c_int64)
c_u32p = C.POINTER(C.c_uint32)
c_f64p = C.POINTER(C.c_double)
vp = C.c_void_p


class Mask(C.Structure):
    _fields_ = [("cond_c_host", vp), ("cond_s_host", vp), ("s_period", C.c_int32)]


class MixcdfArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int64), ("S", C.c_int64), ("C", C.c_int32), ("K", C.c_int32),
        ("z", vp), ("nn_out", vp), ("mask", Mask), ("pad", vp),
        ("scaling_factor", vp), ("mixture_scaling_factor", vp),
        ("reg_max", C.c_float), ("reg_factor", C.c_float), ("training", C.c_int32), ("accumulate", C.c_int32),
        ("params_prebounded", C.c_int32),
        ("z_out", vp), ("ldj", vp), ("reg_ldj", vp), ("status", vp),
        ("next_actnorm_bias", vp), ("next_actnorm_scales", vp), ("next_conv_weight", vp),
        ("nn_compact", C.c_int32),
    ]


class AffineArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int64), ("S", C.c_int64), ("C", C.c_int32),
        ("z", vp), ("nn_out", vp), ("mask", Mask), ("scaling_factor", vp), ("reverse", C.c_int32),
        ("params_prebounded", C.c_int32),
        ("z_out", vp), ("ldj", vp), ("status", vp),
    ]


class ActnormArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int64), ("S", C.c_int64), ("C", C.c_int32),
        ("z", vp), ("bias", vp), ("scales", vp), ("pad", vp), ("length", vp), ("reverse", C.c_int32),
        ("z_out", vp), ("ldj", vp), ("status", vp),
    ]


class ExtActnormArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int64), ("S", C.c_int64), ("C", C.c_int32),
        ("z", vp), ("ext", vp), ("pad", vp), ("reverse", C.c_int32),
        ("z_out", vp), ("ldj", vp), ("status", vp),
    ]


class ActnormInitArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int64), ("S", C.c_int64), ("C", C.c_int32),
        ("x", vp), ("pad", vp), ("workspace", vp), ("bias", vp), ("scales", vp),
    ]


class InvconvBuildArgs(C.Structure):
    _fields_ = [
        ("C", C.c_int32), ("p", vp), ("l", vp), ("u", vp), ("log_s", vp), ("sign_s", vp), ("weight", vp),
        ("w_out", vp), ("w_inv_out", vp), ("sldj_out", vp),
    ]


class InvconvArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int64), ("S", C.c_int64), ("C", C.c_int32),
        ("z", vp), ("weight", vp), ("sldj", vp), ("pad", vp), ("length", vp), ("reverse", C.c_int32),
        ("z_out", vp), ("ldj", vp), ("status", vp),
        ("pre_actnorm_bias", vp), ("pre_actnorm_scales", vp), ("out_mask", vp), ("z_masked_out", vp),
    ]


class CategEncodeArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int64), ("S", C.c_int64), ("V", C.c_int32), ("D", C.c_int32),
        ("tokens", vp), ("u_noise", vp), ("seed", C.c_uint64), ("offset", C.c_uint64),
        ("table", vp), ("category_prior", vp), ("pad", vp), ("beta", C.c_float),
        ("z_out", vp), ("ldj", vp), ("class_prob_log", vp), ("status", vp),
        ("next_actnorm_bias", vp), ("next_actnorm_scales", vp), ("next_conv_weight", vp),
    ]


class CategDecodeArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int64), ("S", C.c_int64), ("V", C.c_int32), ("D", C.c_int32),
        ("z", vp), ("table", vp), ("category_prior", vp), ("tokens_out", vp),
    ]


class LogisticLogprobArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int64), ("S", C.c_int64), ("C", C.c_int32),
        ("x", vp), ("pad", vp), ("mu", C.c_float), ("sigma", C.c_float), ("accumulate", C.c_int32),
        ("out", vp), ("elementwise", vp), ("add", vp), ("total", vp),
    ]


class LogisticSampleArgs(C.Structure):
    _fields_ = [
        ("n", C.c_int64), ("u_noise", vp), ("seed", C.c_uint64), ("offset", C.c_uint64),
        ("mu", C.c_float), ("sigma", C.c_float), ("eps", C.c_float), ("x_out", vp),
    ]


class LdjAxpyArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int64), ("alpha", C.c_float), ("alpha_dev", vp), ("x", vp), ("length", vp), ("y", vp),
    ]


class LinearArgs(C.Structure):
    _fields_ = [
        ("M", C.c_int64), ("N", C.c_int32), ("K", C.c_int32),
        ("x", vp), ("weight", vp), ("bias", vp), ("precision", C.c_int32), ("activation", C.c_int32), ("y", vp),
        ("weight_lo", vp), ("block_n", C.c_int32),
    ]


class LinearBwdArgs(C.Structure):
    _fields_ = [
        ("M", C.c_int64), ("N", C.c_int32), ("K", C.c_int32),
        ("x", vp), ("weight", vp), ("grad_y", vp), ("precision", C.c_int32),
        ("grad_x", vp), ("grad_weight", vp), ("grad_bias", vp), ("weight_lo", vp),
    ]


class CategEncodeBwdArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int64), ("S", C.c_int64), ("V", C.c_int32), ("D", C.c_int32),
        ("tokens", vp), ("z", vp), ("table", vp), ("category_prior", vp), ("pad", vp), ("beta", C.c_float),
        ("grad_z", vp), ("grad_ldj", vp), ("grad_table", vp),
    ]


class LayernormArgs(C.Structure):
    _fields_ = [("M", C.c_int64), ("H", C.c_int32), ("x", vp), ("gamma", vp), ("beta", vp), ("eps", C.c_float), ("y", vp)]


class GraphAttnScoresArgs(C.Structure):
    _fields_ = [
        ("M", C.c_int64), ("E", C.c_int32), ("H", C.c_int32), ("Dh", C.c_int32),
        ("hs", vp), ("hr", vp), ("ld_hs", C.c_int64), ("ld_hr", C.c_int64),
        ("attn_weight", vp), ("score_s", vp), ("score_r", vp),
    ]


class GraphAggregateArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int64), ("N", C.c_int32), ("E", C.c_int32), ("H", C.c_int32), ("Dh", C.c_int32),
        ("adjacency", vp), ("hs", vp), ("hr", vp), ("ld_hs", C.c_int64), ("ld_hr", C.c_int64),
        ("score_s", vp), ("score_r", vp), ("ld_score_s", C.c_int64), ("ld_score_r", C.c_int64), ("num_neighbours", vp),
        ("mode", C.c_int32), ("leaky_slope", C.c_float), ("activation", C.c_int32), ("out", vp),
    ]


class SkipGateArgs(C.Structure):
    _fields_ = [("M", C.c_int64), ("H", C.c_int32), ("config", C.c_int32), ("orig", vp), ("skip", vp), ("out", vp)]


class EdgeAggregateArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int64), ("N", C.c_int32), ("H", C.c_int32), ("Dh", C.c_int32), ("R", C.c_int64),
        ("rev", vp), ("node_val", vp), ("node_q", vp), ("node_k", vp), ("edge_val", vp), ("edge_logit", vp),
        ("ld_node_val", C.c_int64), ("ld_node_q", C.c_int64), ("ld_node_k", C.c_int64), ("ld_edge_val", C.c_int64),
        ("ld_edge_logit", C.c_int64), ("mode", C.c_int32), ("scale", C.c_float), ("out", vp),
    ]


class PairCombineArgs(C.Structure):
    _fields_ = [
        ("R", C.c_int64), ("N", C.c_int32), ("He", C.c_int32),
        ("flat_indices", vp), ("x_indices1", vp), ("x_indices2", vp), ("edge_lin", vp), ("node_lin", vp),
        ("ld_edge", C.c_int64), ("ld_node", C.c_int64), ("activation", C.c_int32), ("out", vp),
    ]


class LinearMixcdfArgs(C.Structure):
    _fields_ = [
        ("mix", MixcdfArgs), ("H", C.c_int32), ("precision", C.c_int32),
        ("features", vp), ("weight", vp), ("bias", vp), ("next_mask", vp), ("z_masked_out", vp),
    ]


class MixcdfBwdArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int64), ("S", C.c_int64), ("C", C.c_int32), ("K", C.c_int32),
        ("z", vp), ("nn_out", vp), ("mask", Mask), ("pad", vp),
        ("scaling_factor", vp), ("mixture_scaling_factor", vp),
        ("reg_max", C.c_float), ("reg_factor", C.c_float), ("training", C.c_int32), ("params_prebounded", C.c_int32),
        ("grad_z_out", vp), ("grad_ldj", vp), ("grad_z", vp), ("grad_nn_out", vp),
        ("grad_scaling_factor", vp), ("grad_mixture_scaling_factor", vp), ("nn_compact", C.c_int32), ("grad_nn_colsum", vp),
        ("proj_weight", vp), ("grad_proj_weight", vp),
    ]


class AffineBwdArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int64), ("S", C.c_int64), ("C", C.c_int32),
        ("z", vp), ("nn_out", vp), ("mask", Mask), ("scaling_factor", vp), ("reverse", C.c_int32),
        ("params_prebounded", C.c_int32),
        ("grad_z_out", vp), ("grad_ldj", vp), ("grad_z", vp), ("grad_nn_out", vp), ("grad_scaling_factor", vp),
    ]


class ActnormBwdArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int64), ("S", C.c_int64), ("C", C.c_int32),
        ("z", vp), ("bias", vp), ("scales", vp), ("pad", vp), ("length", vp), ("reverse", C.c_int32),
        ("grad_z_out", vp), ("grad_ldj", vp), ("grad_z", vp), ("grad_bias", vp), ("grad_scales", vp),
    ]


class ExtActnormBwdArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int64), ("S", C.c_int64), ("C", C.c_int32),
        ("z", vp), ("ext", vp), ("pad", vp), ("reverse", C.c_int32),
        ("grad_z_out", vp), ("grad_ldj", vp), ("grad_z", vp), ("grad_ext", vp),
    ]


class InvconvBwdArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int64), ("S", C.c_int64), ("C", C.c_int32),
        ("z", vp), ("weight", vp), ("pad", vp), ("length", vp), ("reverse", C.c_int32),
        ("grad_z_out", vp), ("grad_ldj", vp), ("grad_z", vp), ("grad_weight", vp), ("grad_sldj", vp),
    ]


class LogisticLogprobBwdArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int64), ("S", C.c_int64), ("C", C.c_int32),
        ("x", vp), ("pad", vp), ("mu", C.c_float), ("sigma", C.c_float),
        ("grad_out", vp), ("grad_elementwise", vp), ("grad_x", vp),
    ]


class SigmoidFlowArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int64), ("n_per_sample", C.c_int64), ("z", vp), ("reverse", C.c_int32), ("alpha", C.c_float),
        ("accumulate", C.c_int32), ("add_tokens", vp), ("z_out", vp), ("ldj", vp), ("ldj_elementwise", vp), ("status", vp),
    ]


class SigmoidFlowBwdArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int64), ("n_per_sample", C.c_int64), ("z", vp), ("reverse", C.c_int32), ("alpha", C.c_float),
        ("grad_z_out", vp), ("grad_ldj", vp), ("grad_ldj_elementwise", vp), ("grad_z", vp),
    ]


class DequantFloorArgs(C.Structure):
    _fields_ = [("n", C.c_int64), ("V", C.c_int32), ("z", vp), ("tokens_out", vp)]


class GeluArgs(C.Structure):
    _fields_ = [("n", C.c_int64), ("x", vp), ("grad_y", vp), ("y", vp)]


class LayernormBwdArgs(C.Structure):
    _fields_ = [("M", C.c_int64), ("H", C.c_int32), ("x", vp), ("gamma", vp), ("eps", C.c_float), ("grad_y", vp),
                ("grad_x", vp), ("grad_gamma", vp), ("grad_beta", vp)]


class SkipGateBwdArgs(C.Structure):
    _fields_ = [("M", C.c_int64), ("H", C.c_int32), ("config", C.c_int32), ("orig", vp), ("skip", vp), ("grad_out", vp),
                ("grad_orig", vp), ("grad_skip", vp)]


class GraphAggregateBwdArgs(C.Structure):
    _fields_ = [("fwd", GraphAggregateArgs), ("grad_out", vp), ("grad_hs", vp), ("grad_hr", vp), ("grad_score_s", vp),
                ("grad_score_r", vp), ("ld_grad_hs", C.c_int64), ("ld_grad_hr", C.c_int64), ("ld_grad_score_s", C.c_int64),
                ("ld_grad_score_r", C.c_int64)]


class EdgeAggregateBwdArgs(C.Structure):
    _fields_ = [("fwd", EdgeAggregateArgs), ("grad_out", vp), ("grad_node_val", vp), ("grad_node_q", vp), ("grad_node_k", vp),
                ("grad_edge_val", vp), ("grad_edge_logit", vp), ("ld_grad_node_val", C.c_int64), ("ld_grad_node_q", C.c_int64),
                ("ld_grad_node_k", C.c_int64), ("ld_grad_edge_val", C.c_int64), ("ld_grad_edge_logit", C.c_int64)]


class PairCombineBwdArgs(C.Structure):
    _fields_ = [("fwd", PairCombineArgs), ("grad_out", vp), ("grad_edge_lin", vp), ("grad_node_lin", vp),
                ("ld_grad_edge", C.c_int64), ("ld_grad_node", C.c_int64)]


# symbol -> argument struct; every entry point is `int f(const Args*, cnf_stream_t)`
ENTRY_POINTS = {
    "cnf_mixcdf_fwd": MixcdfArgs,
    "cnf_mixcdf_inv": MixcdfArgs,
    "cnf_affine_coupling": AffineArgs,
    "cnf_actnorm": ActnormArgs,
    "cnf_ext_actnorm": ExtActnormArgs,
    "cnf_actnorm_data_init": ActnormInitArgs,
    "cnf_invconv_build": InvconvBuildArgs,
    "cnf_invconv_apply": InvconvArgs,
    "cnf_categ_encode": CategEncodeArgs,
    "cnf_categ_decode": CategDecodeArgs,
    "cnf_logistic_logprob": LogisticLogprobArgs,
    "cnf_logistic_sample": LogisticSampleArgs,
    "cnf_ldj_axpy": LdjAxpyArgs,
    "cnf_linear_fwd": LinearArgs,
    "cnf_linear_bwd": LinearBwdArgs,
    "cnf_layernorm": LayernormArgs,
    "cnf_graph_attn_scores": GraphAttnScoresArgs,
    "cnf_graph_aggregate": GraphAggregateArgs,
    "cnf_skip_gate": SkipGateArgs,
    "cnf_edge_aggregate": EdgeAggregateArgs,
    "cnf_pair_combine": PairCombineArgs,
    "cnf_linear_mixcdf_fwd": LinearMixcdfArgs,
    "cnf_linear_mixcdf_inv": LinearMixcdfArgs,
    "cnf_mixcdf_bwd": MixcdfBwdArgs,
    "cnf_affine_coupling_bwd": AffineBwdArgs,
    "cnf_actnorm_bwd": ActnormBwdArgs,
    "cnf_ext_actnorm_bwd": ExtActnormBwdArgs,
    "cnf_invconv_bwd": InvconvBwdArgs,
    "cnf_logistic_logprob_bwd": LogisticLogprobBwdArgs,
    "cnf_categ_encode_bwd": CategEncodeBwdArgs,
    "cnf_sigmoid_flow": SigmoidFlowArgs,
    "cnf_sigmoid_flow_bwd": SigmoidFlowBwdArgs,
    "cnf_dequant_floor": DequantFloorArgs,
    "cnf_gelu": GeluArgs,
    "cnf_layernorm_bwd": LayernormBwdArgs,
    "cnf_skip_gate_bwd": SkipGateBwdArgs,
    "cnf_graph_aggregate_bwd": GraphAggregateBwdArgs,
    "cnf_edge_aggregate_bwd": EdgeAggregateBwdArgs,
    "cnf_pair_combine_bwd": PairCombineBwdArgs,
}
PLAIN_SYMBOLS = ("cnf_last_error_string", "cnf_abi_version", "cnf_built_for_sm", "cnf_mixcdf_fusable", "cnf_mixcdf_path",
                 "cnf_categ_encode_fusable", "cnf_linear_mixcdf_fusable")

ABI_VERSION = 5
_lib = None


class CnfError(RuntimeError):
    """A C-ABI call returned a non-zero status."""

    def __init__(self, fn, code, text):
        super().__init__("%s failed with code %d: %s" % (fn, code, text))
        self.fn, self.code, self.text = fn, code, text


def load():
    """Load ``libcnf_b200.so`` (once) and declare prototypes.  Raises if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "categoricalnf_b200: %s is missing. Build it with `python -m categoricalnf_b200.build` "
            "(needs nvcc); there is no CPU or PyTorch fallback for the hot path." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, struct in ENTRY_POINTS.items():
        fn = getattr(lib, name)
        fn.argtypes = [C.POINTER(struct), vp]
        fn.restype = C.c_int
    lib.cnf_last_error_string.restype = C.c_char_p
    lib.cnf_last_error_string.argtypes = []
    lib.cnf_abi_version.restype = C.c_int
    lib.cnf_built_for_sm.restype = C.c_int
    lib.cnf_mixcdf_fusable.argtypes = [C.POINTER(MixcdfArgs)]
    lib.cnf_mixcdf_fusable.restype = C.c_int
    lib.cnf_mixcdf_path.argtypes = [C.POINTER(MixcdfArgs)]
    lib.cnf_mixcdf_path.restype = C.c_int
    lib.cnf_categ_encode_fusable.argtypes = [C.POINTER(CategEncodeArgs)]
    lib.cnf_categ_encode_fusable.restype = C.c_int
    lib.cnf_linear_mixcdf_fusable.argtypes = [C.POINTER(LinearMixcdfArgs)]
    lib.cnf_linear_mixcdf_fusable.restype = C.c_int
    if lib.cnf_abi_version() != ABI_VERSION:
        raise ImportError("libcnf_b200.so has ABI version %d, the Python binding expects %d - rebuild"
                          % (lib.cnf_abi_version(), ABI_VERSION))
    _lib = lib
    return lib


_entry = {}      # symbol -> bound foreign function (one dict lookup per launch instead of a CDLL attribute walk)


def call(name, args, stream):
    fn = _entry.get(name)
    if fn is None:
        fn = _entry[name] = getattr(load(), name)
    rc = fn(C.byref(args), stream)
    if rc != 0:
        raise CnfError(name, rc, load().cnf_last_error_string().decode("utf-8", "replace"))
