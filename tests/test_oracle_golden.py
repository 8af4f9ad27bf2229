"""Pin the oracle (oracle/cnf_oracle.py) against outputs of the unmodified reference
(tests/golden/*.npz, produced by tests/golden/make_golden.py).  CPU only."""
import pytest
import torch

from conftest import assert_close, load_golden
from oracle import cnf_oracle as O

MIXCDF_CASES = ["mixcdf_lm_small", "mixcdf_lm_padded_sf", "mixcdf_stress", "mixcdf_mol_nodes",
                "mixcdf_mol_edges", "mixcdf_chess", "mixcdf_chess_flip", "mixcdf_flip_k10",
                "mixcdf_ratio_k3"]
TIGHT = dict(rtol=2e-6, atol=2e-6)


def _pad(g):
    if not g.padded:
        return None
    S = g.z.shape[1]
    return (torch.arange(S).view(1, S) < g.length.view(-1, 1)).float().unsqueeze(-1)


@pytest.mark.parametrize("name", MIXCDF_CASES)
def test_mixcdf_forward_and_inverse(name):
    g = load_golden(name)
    m = O.expand_mask(g.mask, g.z)
    kw = dict(pad=_pad(g), reg_max=g.reg_max, reg_factor=g.reg_factor, training=bool(g.training))
    z, ldj, reg = O.mixcdf_coupling(g.z, g.nn_out, m, g.K, g.sf, g.msf, **kw)
    assert_close(z, g.z_fwd, what="z_fwd", **TIGHT)
    assert_close(ldj, g.ldj_fwd, what="ldj_fwd", **TIGHT)
    assert_close(reg, g.reg_ldj, what="reg_ldj", **TIGHT)
    zr, lr, _ = O.mixcdf_coupling(g.z_fwd, g.nn_out, m, g.K, g.sf, g.msf, reverse=True, **kw)
    assert_close(zr, g.z_rev, what="z_rev", **TIGHT)
    assert_close(lr, g.ldj_rev, what="ldj_rev", **TIGHT)
    zs, ls, _ = O.mixcdf_coupling(g.z_lat, g.nn_out, m, g.K, g.sf, g.msf, reverse=True, **kw)
    assert_close(zs, g.z_smp, what="z_smp", **TIGHT)
    assert_close(ls, g.ldj_smp, what="ldj_smp", **TIGHT)


@pytest.mark.parametrize("name", ["mixcdf_tails", "mixcdf_right_tail"])
def test_mixcdf_tails(name):
    g = load_golden(name)
    m = O.expand_mask(g.mask, g.z)
    z, ldj, _ = O.mixcdf_coupling(g.z, g.nn_out, m, g.K, g.sf, g.msf, training=False)
    assert_close(z, g.z_fwd, what="z", **TIGHT)
    assert_close(ldj, g.ldj_fwd, what="ldj", **TIGHT)


def test_mixcdf_inverse_underflow_case():
    """Sampling direction, K = 16: the fixture behind the GPU regression test of the inverse solver."""
    g = load_golden("mixcdf_inv_underflow")
    m = O.expand_mask(g.mask, g.z_lat)
    zs, ls, _ = O.mixcdf_coupling(g.z_lat, g.nn_out, m, g.K, g.sf, g.msf, reverse=True, training=False)
    assert_close(zs, g.z_smp, what="z_smp", **TIGHT)
    assert_close(ls, g.ldj_smp, what="ldj_smp", **TIGHT)


def test_reference_selftest_case():
    """The reference's own __main__ block: forward then reverse reconstructs to 1.2e-7."""
    g = load_golden("mixcdf_selftest")
    m = O.expand_mask(g.mask, g.z)
    z, ldj, _ = O.mixcdf_coupling(g.z, g.nn_out, m, g.K, g.sf, g.msf)
    assert_close(z, g.z_fwd, **TIGHT)
    assert_close(ldj, g.ldj_fwd, **TIGHT)
    zr, lr, _ = O.mixcdf_coupling(z, g.nn_out_rev, m, g.K, g.sf, g.msf, reverse=True)
    assert_close(zr, g.z_rev, **TIGHT)
    assert (zr - g.z).abs().max() < 5e-7
    assert (ldj + lr).abs().max() < 1e-5


def test_autoregressive():
    g = load_golden("autoregressive_mixcdf")
    z, ldj = O.autoregressive_mixcdf(g.z, g.nn_out, g.K, g.sf, g.msf, ldj=g.ldj_in, pad=g.pad)
    assert_close(z, g.z_out, **TIGHT)
    assert_close(ldj, g.ldj_out, **TIGHT)


@pytest.mark.parametrize("name", ["affine_coupling", "affine_coupling_tokens"])
def test_affine(name):
    g = load_golden(name)
    m = O.expand_mask(g.mask, g.z)
    z, ldj = O.affine_coupling(g.z, g.nn_out, m, g.sf, ldj=g.ldj_in)
    assert_close(z, g.z_fwd, **TIGHT)
    assert_close(ldj, g.ldj_fwd, **TIGHT)
    zr, lr = O.affine_coupling(g.z_fwd, g.nn_out, m, g.sf, ldj=g.ldj_fwd, reverse=True)
    assert_close(zr, g.z_rev, **TIGHT)
    assert_close(lr, g.ldj_rev, **TIGHT)


def test_actnorm():
    g = load_golden("actnorm")
    for tag, kw in (("plain", {}), ("len", dict(length=g.length, pad=g.pad)), ("padonly", dict(pad=g.pad))):
        z, ldj = O.actnorm(g.z, g.bias, g.scales, ldj=g["ldj_in_" + tag], **kw)
        assert_close(z, g["z_fwd_" + tag], **TIGHT)
        assert_close(ldj, g["ldj_fwd_" + tag], **TIGHT)
        zr, lr = O.actnorm(g["z_fwd_" + tag], g.bias, g.scales, ldj=g["ldj_fwd_" + tag], reverse=True, **kw)
        assert_close(zr, g["z_rev_" + tag], **TIGHT)
        assert_close(lr, g["ldj_rev_" + tag], **TIGHT)
    b, s = O.actnorm_data_init(g.z, g.pad)
    assert_close(b, g.init_bias_pad, **TIGHT)
    assert_close(s, g.init_scales_pad, **TIGHT)
    b, s = O.actnorm_data_init(g.z)
    assert_close(b, g.init_bias, **TIGHT)
    assert_close(s, g.init_scales, **TIGHT)


def test_ext_actnorm():
    g = load_golden("ext_actnorm")
    out = torch.nn.functional.linear(g.ext, g.weight, g.bias)
    b, s = out.chunk(2, dim=2)
    z, ldj = O.ext_actnorm(g.z, b, s, ldj=g.ldj_in, pad=g.pad)
    assert_close(z, g.z_fwd, **TIGHT)
    assert_close(ldj, g.ldj_fwd, **TIGHT)
    zr, lr = O.ext_actnorm(g.z_fwd, b, s, ldj=g.ldj_fwd, reverse=True, pad=g.pad)
    assert_close(zr, g.z_rev, **TIGHT)
    assert_close(lr, g.ldj_rev, **TIGHT)
    z, ldj = O.ext_actnorm(g.z, b, s, ldj=g.ldj_in)
    assert_close(z, g.z_fwd_nopad, **TIGHT)
    assert_close(ldj, g.ldj_fwd_nopad, **TIGHT)


@pytest.mark.parametrize("C", [2, 6, 16])
def test_invconv(C):
    g = load_golden("invconv")
    t = "_c%d" % C
    w, sldj = O.invconv_weight(g["p" + t], g["l" + t], g["log_s" + t], g["u" + t], g["sign_s" + t])
    assert_close(w, g["w" + t], **TIGHT)
    assert_close(sldj, g["sldj" + t], **TIGHT)
    w_inv = O.invconv_inverse(w)
    assert_close(w_inv, g["w_inv" + t], **TIGHT)
    kw = dict(length=g["length" + t], pad=g["pad" + t])
    z, ldj = O.invconv(g["z" + t], w, sldj, ldj=g["ldj_in" + t], **kw)
    assert_close(z, g["z_fwd" + t], **TIGHT)
    assert_close(ldj, g["ldj_fwd" + t], **TIGHT)
    zr, lr = O.invconv(g["z_fwd" + t], w_inv, sldj, ldj=g["ldj_fwd" + t], reverse=True, **kw)
    assert_close(zr, g["z_rev" + t], **TIGHT)
    assert_close(lr, g["ldj_rev" + t], **TIGHT)
    z, ldj = O.invconv(g["z" + t], w, sldj, ldj=g["ldj_in" + t])
    assert_close(z, g["z_fwd_plain" + t], **TIGHT)
    assert_close(ldj, g["ldj_fwd_plain" + t], **TIGHT)


def test_logistic():
    g = load_golden("logistic")
    assert_close(O.logistic_from_uniform(g.u), g.x, **TIGHT)
    assert_close(O.logistic_log_prob(g.xs), g.log_prob, **TIGHT)


@pytest.mark.parametrize("name", ["encode_lm", "encode_mol_nodes", "encode_mol_edges", "encode_virtual"])
def test_categ_encode_decode(name):
    g = load_golden(name)
    table = O.categ_table(g.embed, g.weight, g.bias)
    pad = g.pad if g.padded else None
    z, ldj, _ = O.categ_encode(g.x, g.u, table, g.category_prior, beta=g.beta, pad=pad)
    assert_close(z, g.z, **TIGHT)
    assert_close(g.ldj_in + ldj, g.ldj, rtol=1e-5, atol=1e-4)
    assert torch.equal(O.categ_decode(g.z, table, g.category_prior), g.x_dec)
    assert torch.equal(O.categ_decode(g.z_rand, table, g.category_prior), g.x_dec_rand)


def test_lm_flow_composition():
    g = load_golden("lm_flow_small")
    enc = dict(table=O.categ_table(g.embed, g.enc_weight, g.enc_bias), prior=g.category_prior)
    blocks = [dict(bias=g["an_bias%d" % i], scales=g["an_scales%d" % i], weight=g["ic_w%d" % i],
                   sldj=g["ic_sldj%d" % i], nn_out=g["nn_out%d" % i], mask=g["mask%d" % i], K=g.K,
                   sf=g["sf%d" % i], msf=g["msf%d" % i]) for i in range(g.NB)]
    z, ldj, logp = O.lm_flow_forward(g.x, g.u, enc, blocks, pad=g.pad, length=g.length)
    assert_close(z, g.z, rtol=1e-5, atol=1e-5)
    assert_close(ldj, g.ldj, rtol=1e-5, atol=1e-4)
    assert_close(logp, g.logp, rtol=1e-5, atol=1e-4)
    # and with the stand-in coupling nets evaluated instead of replaying nn_out
    for i, b in enumerate(blocks):
        w0, b0, w1, b1 = (g["net_%s_%d" % (k, i)] for k in ("w0", "b0", "w1", "b1"))
        b.pop("nn_out")
        b["nn_fn"] = (lambda w0, b0, w1, b1: lambda x: torch.nn.functional.linear(
            torch.nn.functional.gelu(torch.nn.functional.linear(x, w0, b0)), w1, b1))(w0, b0, w1, b1)
    z2, ldj2, _ = O.lm_flow_forward(g.x, g.u, enc, blocks, pad=g.pad, length=g.length)
    assert_close(z2, g.z, rtol=1e-4, atol=1e-4)
    assert_close(ldj2, g.ldj, rtol=1e-4, atol=1e-3)


def _ne_weights(g, which, inverse=False):
    w, sldj = O.invconv_weight(g["ic_%s_p" % which], g["ic_%s_l" % which], g["ic_%s_log_s" % which], g["ic_%s_u" % which],
                               g["ic_%s_sign_s" % which])
    return (O.invconv_inverse(w) if inverse else w), sldj


def test_node_edge_coupling():
    """a14: NodeEdgeCoupling forward (training, regulariser on), eval forward and reverse."""
    g = load_golden("node_edge_coupling")
    args = (g.z_nodes, g.z_edges, g.nn_nodes, g.nn_edges, g.mask_nodes, g.mask_edges, g.Kn, g.Ke, g.sf_nodes, g.sf_edges,
            g.msf_nodes, g.msf_edges)
    kw = dict(ldj=g.ldj_in, pad=g.pad, mask_valid=g.mask_valid, reg_max=3.5, reg_factor=2.0)
    zn, ze, ldj, rn, re = O.node_edge_coupling(*args, training=True, **kw)
    for a, b, w in ((zn, g.cp_zn, "zn"), (ze, g.cp_ze, "ze"), (ldj, g.cp_ldj, "ldj"), (rn, g.cp_reg_nodes, "reg_nodes"),
                    (re, g.cp_reg_edges, "reg_edges")):
        assert_close(a, b, what=w, **TIGHT)
    zn, ze, ldj, _, _ = O.node_edge_coupling(*args, training=False, **kw)
    assert_close(zn, g.cp_zn_eval, what="zn eval", **TIGHT)
    assert_close(ldj, g.cp_ldj_eval, what="ldj eval", **TIGHT)
    zr, er, lr, _, _ = O.node_edge_coupling(g.cp_zn_eval, g.cp_ze_eval, *args[2:], training=False, reverse=True, **kw)
    assert_close(zr, g.cp_zn_rev, what="zn rev", **TIGHT)
    assert_close(er, g.cp_ze_rev, what="ze rev", **TIGHT)
    assert_close(lr, g.cp_ldj_rev, what="ldj rev", **TIGHT)


def test_node_edge_wrapper():
    """NodeEdgeFlowWrapper around ActNorm and InvertibleConv (edge length = number of valid pairs)."""
    g = load_golden("node_edge_coupling")
    kw = dict(length=g.length, pad=g.pad, mask_valid=g.mask_valid)
    zn, ze, ldj = O.node_edge_wrapper(O.actnorm, g.z_nodes, g.z_edges, (g.an_bias_nodes, g.an_scales_nodes),
                                      (g.an_bias_edges, g.an_scales_edges), ldj=g.ldj_in.clone(), **kw)
    for a, b, w in ((zn, g.an_zn, "an zn"), (ze, g.an_ze, "an ze"), (ldj, g.an_ldj, "an ldj")):
        assert_close(a, b, what=w, **TIGHT)
    zi, ei, li = O.node_edge_wrapper(O.invconv, g.an_zn, g.an_ze, _ne_weights(g, "nodes"), _ne_weights(g, "edges"),
                                     ldj=g.an_ldj.clone(), **kw)
    for a, b, w in ((zi, g.ic_zn, "ic zn"), (ei, g.ic_ze, "ic ze"), (li, g.ic_ldj, "ic ldj")):
        assert_close(a, b, what=w, **TIGHT)
    zr, er, lr = O.node_edge_wrapper(O.invconv, g.ic_zn, g.ic_ze, _ne_weights(g, "nodes", True), _ne_weights(g, "edges", True),
                                     ldj=g.ldj_in.clone(), reverse=True, **kw)
    for a, b, w in ((zr, g.ic_zn_rev, "ic zn rev"), (er, g.ic_ze_rev, "ic ze rev"), (lr, g.ic_ldj_rev, "ic ldj rev")):
        assert_close(a, b, what=w, **TIGHT)


# ---------------------------------------------------------------------------------------------------
# graph coupling networks and the graph-colouring flow (BASELINE config 3)
# ---------------------------------------------------------------------------------------------------
def _sd(g):
    return {k[len("sd__"):]: v for k, v in g.items() if k.startswith("sd__")}


@pytest.mark.parametrize("name", ["rgcn_attention", "rgcn_attention_e3", "rgcn_conv", "rgcn_conv_skip0"])
def test_rgcn_net(name):
    from oracle import graph_oracle as GO
    g = load_golden(name)
    kw = dict(num_edges=g.num_edges, num_layers=g.layers, attention=bool(g.attention), skip_config=g.skip_config,
              max_neighbours=g.max_neighbours)
    out = GO.rgcn_net(_sd(g), g.x, g.adjacency, **kw)
    assert_close(out, g.out, rtol=1e-5, atol=2e-6, what="RGCNNet")
    out = GO.rgcn_net(_sd(g), g.x, g.adjacency, pad=g.pad, **kw)
    assert_close(out, g.out_pad, rtol=1e-5, atol=2e-6, what="RGCNNet padded")


def test_graph_node_flow():
    from oracle import graph_oracle as GO
    g = load_golden("graph_node_flow")
    kw = dict(num_flows=2, num_layers=2, num_mixtures=8)
    z, ldj = GO.graph_node_flow(_sd(g), g.x, g.adjacency, g.length, g.u, **kw)
    assert_close(z, g.z, rtol=1e-5, atol=5e-6, what="z")
    assert_close(ldj, g.ldj, rtol=1e-5, atol=5e-5, what="ldj")
    zr, lr = GO.graph_node_flow(_sd(g), g.x, g.adjacency, g.length, g.u, reverse_z=g.z, **kw)
    assert_close(zr, g.z_rev, rtol=1e-4, atol=2e-5, what="z reverse")
    assert_close(lr, g.ldj_rev, rtol=1e-4, atol=2e-4, what="ldj reverse")


@pytest.mark.parametrize("name", ["edge_gnn_attn_sparse", "edge_gnn_attn_dense", "edge_gnn_qkv_dense", "edge_gnn_qkv_sparse"])
def test_edge_gnn(name):
    from oracle import graph_oracle as GO
    g = load_golden(name)
    binary = (g.adjacency > 0).long() if g.sparse else None
    nodes, edges = GO.edge_gnn(_sd(g), g.z_nodes, g.z_edges, (g.x_indices1, g.x_indices2), g.mask_valid, num_layers=g.layers,
                               qkv=bool(g.qkv), pad=g.pad, binary_adjacency=binary, max_neighbours=g.max_neighbours)
    assert_close(nodes, g.nodes_out, rtol=1e-5, atol=2e-6, what="nodes_out")
    assert_close(edges, g.edges_out, rtol=1e-5, atol=2e-6, what="edges_out")


def test_encoding_with_linear_flows_and_decoder():
    """BASELINE config 1's encoder (4 x [ExtActNorm, InvConv, affine coupling]) and DecoderLinear."""
    g = load_golden("encoding_variants")
    sd = {k[len("sd_flows__"):]: v for k, v in g.items() if k.startswith("sd_flows__")}
    z, ldj = O.categ_encode_flows(sd, g.x, g.flows_u, num_flows=4, beta=0.8, pad=g.pad)
    assert_close(z, g.flows_z, rtol=1e-5, atol=2e-6, what="z")
    assert_close(ldj, g.flows_ldj, rtol=1e-5, atol=2e-5, what="ldj")
    assert torch.equal(O.categ_decode_flows(sd, g.flows_z, num_flows=4), g.flows_x_dec)
    sdd = {k[len("sd_dec__"):]: v for k, v in g.items() if k.startswith("sd_dec__")}
    assert_close(O.decoder_linear(sdd, g.dec_z), g.dec_log_probs, rtol=1e-5, atol=2e-6, what="decoder log-probs")


def test_sigmoid_flow_and_variational_dequantization():
    """SigmoidFlow both directions (summed and element-wise ldj) and VariationalDequantization with 4 flows."""
    g = load_golden("dequantization")
    z, ldj = O.sigmoid_flow(g.sig_in, g.ldj0)
    assert_close(z, g.sig_z, rtol=1e-6, atol=1e-7, what="sigmoid z")
    assert_close(ldj, g.sig_ldj, rtol=1e-6, atol=1e-5, what="sigmoid ldj")
    _, elem = O.sigmoid_flow(g.sig_in, reverse=True, reverse_layer=True, sum_ldj=False)
    assert_close(elem, g.sig_elem, rtol=1e-6, atol=1e-6, what="sigmoid element ldj")
    z, ldj = O.sigmoid_flow(g.logit_in, g.ldj0, reverse=True)
    assert_close(z, g.logit_z, rtol=1e-6, atol=1e-6, what="logit z")
    assert_close(ldj, g.logit_ldj, rtol=1e-6, atol=1e-5, what="logit ldj")
    _, elem = O.sigmoid_flow(g.logit_in, reverse_layer=True, sum_ldj=False)
    assert_close(elem, g.logit_elem, rtol=1e-6, atol=1e-6, what="logit element ldj")
    sd = {k[len("sd__"):]: v for k, v in g.items() if k.startswith("sd__")}
    z_cont, ldj = O.variational_dequantization(sd, g.x, g.u, g.num_flows)
    assert_close(z_cont, g.z_cont, rtol=1e-5, atol=2e-6, what="dequantised z")
    assert_close(ldj, g.ldj, rtol=1e-5, atol=2e-5, what="dequantisation ldj")
    assert torch.equal(O.dequantization_reverse(g.z_cont, g.V), g.x_rec)
    assert ((z_cont.squeeze(-1) - g.x.float()) >= 0).all() and ((z_cont.squeeze(-1) - g.x.float()) <= 1).all()
