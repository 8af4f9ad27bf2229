"""The C-ABI library builds, loads and exports every symbol include/cnf_b200.h declares (CPU only:
no kernel is launched)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "cnf_b200.h")).read()
    return sorted(set(re.findall(r"CNF_API\s+[\w\s\*]*?\b(cnf_\w+)\s*\(", text)))


def test_header_declares_entry_points():
    syms = declared_symbols()
    assert "cnf_mixcdf_fwd" in syms and "cnf_mixcdf_inv" in syms and "cnf_last_error_string" in syms
    assert len(syms) >= 16


def test_library_exports_every_declared_symbol():
    from categoricalnf_b200 import build
    path = build.build()
    lib = ctypes.CDLL(path)
    for s in declared_symbols():
        assert hasattr(lib, s), "libcnf_b200.so does not export %s" % s


def test_python_binding_matches_header():
    from categoricalnf_b200 import _lib
    syms = set(declared_symbols())
    bound = set(_lib.ENTRY_POINTS) | set(_lib.PLAIN_SYMBOLS)
    assert bound == syms, "binding and header disagree: %s" % sorted(bound ^ syms)
    lib = _lib.load()
    assert lib.cnf_abi_version() == _lib.ABI_VERSION
    assert lib.cnf_built_for_sm() == 100


def test_struct_layout_matches_c():
    """sizeof() of every ctypes struct equals the C compiler's (guards against field drift)."""
    import subprocess
    import tempfile
    from categoricalnf_b200 import _lib
    names = {"cnf_mask": _lib.Mask, "cnf_mixcdf_args": _lib.MixcdfArgs, "cnf_affine_args": _lib.AffineArgs,
             "cnf_actnorm_args": _lib.ActnormArgs, "cnf_ext_actnorm_args": _lib.ExtActnormArgs,
             "cnf_actnorm_init_args": _lib.ActnormInitArgs, "cnf_invconv_build_args": _lib.InvconvBuildArgs,
             "cnf_invconv_args": _lib.InvconvArgs, "cnf_categ_encode_args": _lib.CategEncodeArgs,
             "cnf_categ_decode_args": _lib.CategDecodeArgs, "cnf_logistic_logprob_args": _lib.LogisticLogprobArgs,
             "cnf_logistic_sample_args": _lib.LogisticSampleArgs, "cnf_ldj_axpy_args": _lib.LdjAxpyArgs, "cnf_linear_args": _lib.LinearArgs,
             "cnf_linear_bwd_args": _lib.LinearBwdArgs, "cnf_layernorm_args": _lib.LayernormArgs,
             "cnf_graph_attn_scores_args": _lib.GraphAttnScoresArgs, "cnf_graph_aggregate_args": _lib.GraphAggregateArgs,
             "cnf_skip_gate_args": _lib.SkipGateArgs, "cnf_edge_aggregate_args": _lib.EdgeAggregateArgs,
             "cnf_pair_combine_args": _lib.PairCombineArgs, "cnf_categ_encode_bwd_args": _lib.CategEncodeBwdArgs,
             "cnf_linear_mixcdf_args": _lib.LinearMixcdfArgs, "cnf_mixcdf_bwd_args": _lib.MixcdfBwdArgs,
             "cnf_affine_bwd_args": _lib.AffineBwdArgs, "cnf_actnorm_bwd_args": _lib.ActnormBwdArgs,
             "cnf_ext_actnorm_bwd_args": _lib.ExtActnormBwdArgs, "cnf_invconv_bwd_args": _lib.InvconvBwdArgs,
             "cnf_logistic_logprob_bwd_args": _lib.LogisticLogprobBwdArgs}
    prog = '#include <stdio.h>\n#include "cnf_b200.h"\nint main(void){\n'
    for n in names:
        prog += '  printf("%s %%zu\\n", sizeof(%s));\n' % (n, n)
    prog += "  return 0;}\n"
    with tempfile.TemporaryDirectory() as d:
        src, exe = os.path.join(d, "s.c"), os.path.join(d, "s")
        open(src, "w").write(prog)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        out = subprocess.check_output([exe], text=True)
    for line in out.strip().splitlines():
        n, size = line.split()
        assert ctypes.sizeof(names[n]) == int(size), "%s: ctypes %d vs C %s" % (n, ctypes.sizeof(names[n]), size)


def test_cpu_tensors_are_rejected():
    import torch
    from categoricalnf_b200 import ops
    z = torch.zeros(2, 3, 4)
    with pytest.raises(RuntimeError, match="CUDA only|no CPU fallback"):
        ops.mixcdf(z, torch.zeros(2, 3, 4 * 26), 8, mask_c=[1, 1, 0, 0])
    with pytest.raises(RuntimeError, match="CUDA only|no CPU fallback"):
        ops.actnorm(z, torch.zeros(4), torch.zeros(4))
