"""GPU parity of the GraphCNF node+edge block: the drop-in ``NodeEdgeCoupling`` /
``NodeEdgeFlowWrapper`` modules against the golden outputs of the unmodified reference classes
(experiments/molecule_generation/graph_node_edge_coupling.py) and against the oracle at the Zinc shape."""
import pytest
import torch
import torch.nn as nn

from conftest import assert_close, load_golden
from oracle import cnf_oracle as O

pytestmark = pytest.mark.gpu


def ldj_close(a, b, what):
    assert_close(a, b, rtol=1e-4, atol=2e-4, what=what)


class _Preset(nn.Module):
    """Stand-in Edge-GNN returning preset outputs (the fixture stores them as explicit inputs)."""

    def __init__(self, nodes, edges):
        super().__init__()
        self.nodes, self.edges = nodes, edges

    def forward(self, z_nodes, z_edges, **kwargs):
        return self.nodes, self.edges


def _coupling(g, dev="cuda"):
    from categoricalnf_b200.layers.flows.node_edge_coupling import NodeEdgeCoupling
    cp = NodeEdgeCoupling(c_in_nodes=g.Cn, c_in_edges=g.Ce, mask_nodes=g.mask_nodes, mask_edges=g.mask_edges,
                          num_mixtures_nodes=g.Kn, num_mixtures_edges=g.Ke,
                          model_func=lambda c_out_nodes, c_out_edges: _Preset(g.nn_nodes.to(dev), g.nn_edges.to(dev)),
                          regularizer_max=3.5, regularizer_factor=2).to(dev)
    cp.scaling_factor_nodes.data = g.sf_nodes.to(dev)
    cp.scaling_factor_edges.data = g.sf_edges.to(dev)
    cp.mixture_scaling_factor_nodes.data = g.msf_nodes.to(dev)
    cp.mixture_scaling_factor_edges.data = g.msf_edges.to(dev)
    return cp


def test_node_edge_coupling_golden():
    g = load_golden("node_edge_coupling")
    cp = _coupling(g)
    kw = dict(length=g.length.cuda(), channel_padding_mask=g.pad.cuda(), mask_valid=g.mask_valid.cuda())
    with torch.no_grad():
        cp.train()
        zn, ze, ldj, detail = cp(g.z_nodes.cuda(), g.z_edges.cuda(), ldj=g.ldj_in.cuda(), **kw)
        assert_close(zn, g.cp_zn, what="nodes")
        assert_close(ze, g.cp_ze, what="edges")
        ldj_close(ldj, g.cp_ldj, "ldj")
        ldj_close(detail["regularizer_nodes_ldj"], g.cp_reg_nodes, "reg nodes")
        ldj_close(detail["regularizer_edges_ldj"], g.cp_reg_edges, "reg edges")
        cp.eval()
        zn, ze, ldj, _ = cp(g.z_nodes.cuda(), g.z_edges.cuda(), ldj=g.ldj_in.cuda(), **kw)
        assert_close(zn, g.cp_zn_eval, what="nodes eval")
        ldj_close(ldj, g.cp_ldj_eval, "ldj eval")
        zr, er, lr, detail = cp(g.cp_zn_eval.cuda(), g.cp_ze_eval.cuda(), ldj=g.ldj_in.cuda(), reverse=True, **kw)
        assert_close(zr, g.cp_zn_rev, what="nodes reverse")
        assert_close(er, g.cp_ze_rev, what="edges reverse")
        ldj_close(lr, g.cp_ldj_rev, "ldj reverse")
        assert "regularizer_nodes_ldj" not in detail


def test_node_edge_wrapper_golden():
    from categoricalnf_b200.layers.flows import ActNormFlow, InvertibleConv
    from categoricalnf_b200.layers.flows.node_edge_coupling import NodeEdgeFlowWrapper
    g = load_golden("node_edge_coupling")
    an = NodeEdgeFlowWrapper(ActNormFlow(g.Cn), ActNormFlow(g.Ce))
    ic = NodeEdgeFlowWrapper(InvertibleConv(g.Cn), InvertibleConv(g.Ce))
    an.node_flow.bias.data, an.node_flow.scales.data = g.an_bias_nodes, g.an_scales_nodes
    an.edge_flow.bias.data, an.edge_flow.scales.data = g.an_bias_edges, g.an_scales_edges
    for flow, which in ((ic.node_flow, "nodes"), (ic.edge_flow, "edges")):
        flow.load_state_dict({k[len("ic_%s_" % which):]: v for k, v in g.items() if k.startswith("ic_%s_" % which)})
    an, ic = an.cuda().eval(), ic.cuda().eval()
    kw = dict(length=g.length.cuda(), channel_padding_mask=g.pad.cuda(), mask_valid=g.mask_valid.cuda())
    with torch.no_grad():
        zn, ze, ldj = an(g.z_nodes.cuda(), g.z_edges.cuda(), ldj=g.ldj_in.cuda(), **kw)
        assert_close(zn, g.an_zn, what="actnorm nodes")
        assert_close(ze, g.an_ze, what="actnorm edges")
        ldj_close(ldj, g.an_ldj, "actnorm ldj")
        zi, ei, li = ic(zn, ze, ldj=ldj.clone(), **kw)
        assert_close(zi, g.ic_zn, what="conv nodes")
        assert_close(ei, g.ic_ze, what="conv edges")
        ldj_close(li, g.ic_ldj, "conv ldj")
        zr, er, lr = ic(g.ic_zn.cuda(), g.ic_ze.cuda(), ldj=g.ldj_in.cuda(), reverse=True, **kw)
        assert_close(zr, g.ic_zn_rev, what="conv nodes reverse")
        assert_close(er, g.ic_ze_rev, what="conv edges reverse")
        ldj_close(lr, g.ic_ldj_rev, "conv ldj reverse")


def test_node_edge_coupling_zinc_shape_vs_oracle():
    """BASELINE config 4 per-GPU shape: B 64, N 38 -> 703 pairs, nodes C 6 / K 16, edges C 2 / K 8."""
    from categoricalnf_b200.layers.flows import CouplingLayer
    from categoricalnf_b200.layers.flows.node_edge_coupling import NodeEdgeCoupling
    gen = torch.Generator().manual_seed(5)
    B, N, Cn, Ce, Kn, Ke = 64, 38, 6, 2, 16, 8
    P = N * (N - 1) // 2
    length = torch.randint(20, N + 1, (B,), generator=gen)
    pad = (torch.arange(N)[None, :] < length[:, None]).float().unsqueeze(-1)
    idx = torch.tensor([(i, j) for i in range(N) for j in range(i + 1, N)])
    mask_valid = ((idx[None, :, 0] < length[:, None]) & (idx[None, :, 1] < length[:, None])).float()
    z_nodes = torch.randn(B, N, Cn, generator=gen) * pad
    z_edges = torch.randn(B, P, Ce, generator=gen) * mask_valid.unsqueeze(-1)
    nn_nodes = torch.randn(B, N, Cn * (2 + 3 * Kn), generator=gen) * 0.5
    nn_edges = torch.randn(B, P, Ce * (2 + 3 * Ke), generator=gen) * 0.5
    mn, me = CouplingLayer.create_channel_mask(Cn), CouplingLayer.create_channel_mask(Ce)
    cp = NodeEdgeCoupling(Cn, Ce, mn, me, Kn, Ke, lambda c_out_nodes, c_out_edges: _Preset(nn_nodes.cuda(), nn_edges.cuda()),
                          regularizer_max=3.5, regularizer_factor=2).cuda().eval()
    zeros = lambda *s: torch.zeros(*s)
    ref = O.node_edge_coupling(z_nodes, z_edges, nn_nodes, nn_edges, mn, me, Kn, Ke, zeros(Cn), zeros(Ce), zeros(Cn, Kn),
                               zeros(Ce, Ke), pad=pad, mask_valid=mask_valid, reg_max=3.5, reg_factor=2.0, training=False)
    kw = dict(length=length.cuda(), channel_padding_mask=pad.cuda(), mask_valid=mask_valid.cuda())
    with torch.no_grad():
        zn, ze, ldj, _ = cp(z_nodes.cuda(), z_edges.cuda(), **kw)
        assert_close(zn, ref[0], what="nodes")
        assert_close(ze, ref[1], what="edges")
        ldj_close(ldj, ref[2], "ldj")
        zr, er, lr, _ = cp(zn, ze, reverse=True, **kw)
        assert_close(zr, z_nodes, rtol=1e-4, atol=1e-4, what="nodes round trip")
        assert_close(er, z_edges, rtol=1e-4, atol=1e-4, what="edges round trip")
        assert_close(lr, -ref[2], rtol=1e-4, atol=1e-2, what="ldj antisymmetry")


def test_static_entry_points_as_called_by_reference_node_edge_coupling():
    """The UNMODIFIED reference NodeEdgeCoupling._run_mixture_layer (graph_node_edge_coupling.py:112-140) calls the
    static ``get_mixt_params`` / ``run_with_params`` with a [1,1,C] mask, a float64 ``orig_z`` and reduces the
    regulariser over dims 1..; replay that call sequence against the drop-in class."""
    from categoricalnf_b200.layers.flows import MixtureCDFCoupling
    g = load_golden("node_edge_coupling")
    pad = g.pad.cuda()
    nn_out = g.nn_nodes.cuda() * pad
    mask = g.mask_nodes.cuda()[None, :1, :]
    p = MixtureCDFCoupling.get_mixt_params(nn_out=nn_out, mask=mask, num_mixtures=g.Kn, scaling_factor=g.sf_nodes.cuda(),
                                           mixture_scaling_factor=g.msf_nodes.cuda())
    t, log_s, log_pi, mixt_t, mixt_log_s = p
    z_out, ldj, reg = MixtureCDFCoupling.run_with_params(orig_z=g.z_nodes.cuda().double(), t=t, log_s=log_s, log_pi=log_pi,
                                                         mixt_t=mixt_t, mixt_log_s=mixt_log_s, reverse=False, is_training=True,
                                                         reg_max=3.5, reg_factor=2, mask=mask, channel_padding_mask=pad,
                                                         return_reg_ldj=True)
    reg = reg.float().sum(dim=[i for i in range(1, len(reg.shape))])
    assert reg.shape == (g.B,)
    assert_close(z_out.float() * pad, g.cp_zn, what="nodes")
    ldj_close(reg, g.cp_reg_nodes, "reg nodes")
    # explicit-tensor form (parameters materialised by the caller, as upstream's five float64 tensors)
    t, log_s, log_pi, mu, mls = (v.double() for v in p[0].materialize())
    z2, ldj2 = MixtureCDFCoupling.run_with_params(g.z_nodes.cuda().double(), t, log_s, log_pi, mu, mls, mask=mask,
                                                  channel_padding_mask=pad, is_training=False)
    assert_close(z2.float() * pad, g.cp_zn_eval, what="nodes (explicit parameters)")
