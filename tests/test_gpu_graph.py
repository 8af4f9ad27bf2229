"""GPU parity of the graph coupling networks and the graph-colouring flow (BASELINE config 3):
glue kernels against plain formulas, ``RGCNNet`` (attention and plain relational convolution) and ``GraphNodeFlow``
against the golden outputs of the unmodified reference (state dicts loaded by name, strict), the differentiable
training path against the kernel path and the oracle's autograd, and size-independent properties at the
config's full size (B 1024, N 20, hidden 384).

Tolerance: networks |a-b| <= 1e-4 |b| + 2e-5 (3xTF32 projections, fp32 glue); flow outputs as everywhere:
|a-b| <= 1e-4 |b| + 1e-5 on z, 1e-4 relative (+2e-4) on ldj."""
import pytest
import torch

from conftest import assert_close, load_golden
from oracle import graph_oracle as GO

pytestmark = pytest.mark.gpu


def _sd(g):
    return {k[len("sd__"):]: v for k, v in g.items() if k.startswith("sd__")}


def _graphs(gen, B, N, E, p=0.25):
    length = torch.randint(max(2, N // 2), N + 1, (B,), generator=gen)
    up = (torch.rand(B, N, N, generator=gen) < p).long() * torch.randint(1, E + 1, (B, N, N), generator=gen)
    up = torch.triu(up, diagonal=1)
    adj = up + up.transpose(1, 2)
    valid = torch.arange(N)[None, :] < length[:, None]
    return adj * (valid[:, :, None] & valid[:, None, :]).long(), length


# ---- kernels --------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,H", [(37, 32), (1000, 384), (5, 30), (20480, 96)])
def test_layernorm(M, H):
    from categoricalnf_b200 import ops
    gen = torch.Generator().manual_seed(M + H)
    x = torch.randn(M, H, generator=gen) * 2 + 0.5
    w, b = torch.randn(H, generator=gen), torch.randn(H, generator=gen)
    y = ops.layernorm(x.cuda(), w.cuda(), b.cuda())
    ref = torch.nn.functional.layer_norm(x.double(), (H,), w.double(), b.double(), 1e-5)
    assert_close(y, ref, rtol=1e-5, atol=1e-5, what="layernorm")


@pytest.mark.parametrize("config", [0, 1, 2])
def test_skip_gate(config):
    from categoricalnf_b200 import ops
    gen = torch.Generator().manual_seed(config)
    orig = torch.randn(6, 11, 40, generator=gen)
    s = torch.randn(6, 11, 40 if config == 0 else 80, generator=gen)
    out = ops.skip_gate(orig.cuda(), s.cuda(), config)
    if config == 0:
        ref = orig + s
    else:
        val, gate = s.double().chunk(2, dim=-1)
        gate = torch.sigmoid(gate)
        ref = orig + val * gate if config == 1 else orig * (1 - gate) + val * gate
    assert_close(out, ref, rtol=1e-5, atol=2e-6, what="skip gate")


@pytest.mark.parametrize("B,N,E,H,Dh", [(3, 9, 1, 4, 16), (2, 20, 3, 4, 12), (5, 38, 2, 2, 7), (4, 70, 1, 8, 8)])
def test_attention_aggregate_vs_dense(B, N, E, H, Dh):
    """cnf_graph_attn_scores + cnf_graph_aggregate (mode 1) against a dense masked softmax in float64."""
    from categoricalnf_b200 import ops
    gen = torch.Generator().manual_seed(B * N + E)
    adj, _ = _graphs(gen, B, N, E)
    hs = torch.randn(B, N, H * Dh, generator=gen)
    hr = torch.randn(B, N, (E + 1) * H * Dh, generator=gen)
    aw = torch.randn(H, 2, Dh, generator=gen) * 0.5
    out = ops.graph_attention_aggregate(hs.cuda(), hr.cuda(), aw.cuda(), adj.cuda(), E)
    hs4, hr5 = hs.double().view(B, N, H, Dh), hr.double().view(B, N, E + 1, H, Dh)
    s_s = (hs4 * aw[:, 0].double()).sum(-1)
    s_r = (hr5 * aw[:, 1].double()).sum(-1)
    etype = torch.where(torch.eye(N, dtype=torch.bool)[None], torch.full_like(adj, E + 1), adj)
    ref = torch.zeros(B, N, H, Dh, dtype=torch.float64)
    for b in range(B):
        for i in range(N):
            js = torch.nonzero(etype[b, i] > 0).flatten()
            es = etype[b, i, js] - 1
            logit = torch.nn.functional.leaky_relu(s_s[b, i][None, :] + s_r[b, js, es], 0.2)       # [n,H]
            p = torch.softmax(logit, dim=0)
            ref[b, i] = (p[..., None] * hr5[b, js, es]).sum(0)
    assert_close(out, ref.view(B, N, H * Dh), rtol=1e-5, atol=5e-6, what="attention aggregate")
    # column slices of one wider matrix (the fused hs|hr projection output) and the GELU epilogue
    wide = torch.cat([hs, hr], dim=-1).reshape(B * N, -1).cuda()
    out2 = ops.graph_attention_aggregate(wide[:, :H * Dh].unflatten(0, (B, N)), wide[:, H * Dh:].unflatten(0, (B, N)),
                                         aw.cuda(), adj.cuda(), E, activation="gelu")
    assert_close(out2, torch.nn.functional.gelu(ref.view(B, N, H * Dh)), rtol=1e-5, atol=5e-6, what="sliced + gelu")


@pytest.mark.parametrize("B,N,E,C", [(3, 12, 3, 32), (2, 38, 3, 20), (4, 9, 1, 7)])
def test_mean_aggregate_vs_dense(B, N, E, C):
    from categoricalnf_b200 import ops
    gen = torch.Generator().manual_seed(B + N + E + C)
    adj, _ = _graphs(gen, B, N, E, p=0.4)
    hs = torch.randn(B, N, C, generator=gen)
    hr = torch.randn(B, N, E * C, generator=gen)
    nn_ = (adj > 0).sum(dim=1).float().clamp(max=4)
    out = ops.graph_mean_aggregate(hs.cuda(), hr.cuda(), adj.cuda(), E, nn_.cuda())
    onehot = torch.nn.functional.one_hot(adj, E + 1)[..., 1:].double()
    ref = hs.double() + torch.einsum("bjie,bjec->bic", onehot, hr.double().view(B, N, E, C)) / nn_.double().unsqueeze(-1).clamp(min=1e-5)
    assert_close(out, ref, rtol=1e-5, atol=5e-6, what="mean aggregate")
    out = ops.graph_mean_aggregate(hs.cuda(), hr.cuda(), adj.cuda(), E, None)
    cnt = (adj > 0).sum(dim=1).double()
    ref = hs.double() + torch.einsum("bjie,bjec->bic", onehot, hr.double().view(B, N, E, C)) / cnt.unsqueeze(-1).clamp(min=1e-5)
    assert_close(out, ref, rtol=1e-5, atol=5e-6, what="mean aggregate, counted neighbours")


# ---- networks -------------------------------------------------------------------------------------------------------
def _build_rgcn(g):
    from categoricalnf_b200.layers.networks import RGCNNet, RelationGraphAttention, RelationGraphConv
    net = RGCNNet(c_in=g.c_in, c_out=g.c_out, num_edges=g.num_edges, num_layers=g.layers, hidden_size=g.hidden,
                  skip_config=g.skip_config, max_neighbours=g.max_neighbours,
                  rgc_layer_fun=RelationGraphAttention if g.attention else RelationGraphConv)
    net.load_state_dict(_sd(g), strict=True)          # the reference's parameter names
    return net.cuda().eval()


@pytest.mark.parametrize("name", ["rgcn_attention", "rgcn_attention_e3", "rgcn_conv", "rgcn_conv_skip0"])
def test_rgcn_net_golden(name):
    g = load_golden(name)
    net = _build_rgcn(g)
    with torch.no_grad():
        out = net(g.x.cuda(), adjacency=g.adjacency.cuda())
        out_pad = net(g.x.cuda(), adjacency=g.adjacency.cuda(), channel_padding_mask=g.pad.cuda())
    assert_close(out, g.out, rtol=1e-4, atol=2e-5, what="RGCNNet")
    assert_close(out_pad, g.out_pad, rtol=1e-4, atol=2e-5, what="RGCNNet padded")
    # the reference's layers take the one-hot adjacency: same result through that form
    layer = net.layers[0][0]
    h = torch.randn(g.x.shape[0], g.x.shape[1], g.hidden, generator=torch.Generator().manual_seed(1)).cuda()
    onehot = torch.nn.functional.one_hot(g.adjacency, g.num_edges + 1)[..., 1:].float().cuda()
    with torch.no_grad():
        assert_close(layer(h, adjacency=onehot), layer(h, adjacency=g.adjacency.cuda()), rtol=0, atol=0, what="one-hot adjacency")


@pytest.mark.parametrize("name", ["rgcn_attention_e3", "rgcn_conv"])
def test_rgcn_training_path_and_gradients(name):
    """With grad enabled the glue runs as differentiable dense torch ops around the tensor-core projections: same
    output as the kernel path, gradients equal to autograd through the CPU oracle."""
    g = load_golden(name)
    net = _build_rgcn(g).train()
    sd = {k: v.clone().requires_grad_(True) for k, v in _sd(g).items()}
    x_ref = g.x.clone().requires_grad_(True)
    out_ref = GO.rgcn_net(sd, x_ref, g.adjacency, num_edges=g.num_edges, num_layers=g.layers, attention=bool(g.attention),
                          skip_config=g.skip_config, max_neighbours=g.max_neighbours)
    wgt = torch.randn(out_ref.shape, generator=torch.Generator().manual_seed(2))
    (out_ref * wgt).sum().backward()
    x = g.x.cuda().requires_grad_(True)
    out = net(x, adjacency=g.adjacency.cuda())
    assert_close(out, g.out, rtol=1e-4, atol=2e-5, what="training-path output")
    (out * wgt.cuda()).sum().backward()
    assert_close(x.grad, x_ref.grad, rtol=1e-3, atol=2e-5, what="grad x")
    for k, p in net.named_parameters():
        assert p.grad is not None, k
        scale = float(sd[k].grad.abs().max()) + 1e-6
        assert_close(p.grad / scale, sd[k].grad / scale, rtol=1e-3, atol=2e-4, what="grad " + k)


def _build_flow(g, hidden=32, layers=2, flows=2, mixtures=8):
    from categoricalnf_b200.experiments.graph_coloring import GraphNodeFlow

    class _Dataset:
        @staticmethod
        def num_node_types():
            return 3

    params = {"categ_encoding": {"use_dequantization": False, "use_variational": False, "use_decoder": False, "num_dimensions": 2,
                                 "flow_config": {"num_flows": 0, "hidden_layers": 2, "hidden_size": 128},
                                 "decoder_config": {"num_layers": 1, "hidden_size": 64}},
              "coupling_num_flows": flows, "coupling_hidden_size": hidden, "coupling_hidden_layers": layers,
              "coupling_num_mixtures": mixtures, "coupling_mask_ratio": 0.5, "coupling_dropout": 0.0}
    model = GraphNodeFlow(params, _Dataset)
    if g is not None:
        model.load_state_dict(_sd(g), strict=True)
    return model.cuda().eval()


def test_graph_node_flow_golden():
    g = load_golden("graph_node_flow")
    model = _build_flow(g)
    with torch.no_grad():
        z, ldj = model(g.x.cuda(), adjacency=g.adjacency.cuda(), length=g.length.cuda(), u_noise=g.u.cuda())
        assert_close(z, g.z, what="z")
        assert_close(ldj, g.ldj, rtol=1e-4, atol=2e-4, what="ldj")
        # reverse pass of the continuous layers, then decoding
        N = g.x.shape[1]
        from categoricalnf_b200.experiments.graph_coloring.graph_node_flow import length_masks
        kpm, cpm = length_masks(g.length.cuda(), N)
        kw = dict(adjacency=g.adjacency.cuda(), length=g.length.cuda(), channel_padding_mask=cpm, src_key_padding_mask=kpm)
        z_rev, ldj_rev = g.z.cuda(), torch.zeros(g.x.shape[0], device="cuda")
        for layer in reversed(list(model.flow_layers)[1:]):
            res = layer(z_rev, reverse=True, **kw)
            z_rev, ldj_rev = res[0], ldj_rev + res[1]
        assert_close(z_rev, g.z_rev, rtol=1e-4, atol=2e-5, what="z reverse")
        assert_close(ldj_rev, g.ldj_rev, rtol=1e-4, atol=2e-4, what="ldj reverse")
        x_dec = model.node_embed_flow(g.z_rev.cuda(), reverse=True, **kw)[0]
        valid = cpm.squeeze(-1).bool().cpu()
        assert torch.equal(x_dec.cpu()[valid], g.x_dec[valid])


def test_graph_node_flow_full_size_properties():
    """BASELINE config 3 size: B 1024, N 20, hidden 384, 4 layers, 8 flows, 8 mixtures - data-dependent init,
    forward, batch-split consistency of the log-likelihood, forward -> reverse round trip."""
    gen = torch.Generator().manual_seed(0)
    torch.manual_seed(0)
    B, N = 1024, 20
    model = _build_flow(None, hidden=384, layers=4, flows=8, mixtures=8)
    adj, length = _graphs(gen, B, N, 1, p=0.15)
    x = torch.randint(0, 3, (B, N), generator=gen)
    xc, ac, lc = x.cuda(), adj.cuda(), length.cuda()
    with torch.no_grad():
        model.initialize_data_dependent([(xc[:256], {"adjacency": ac[:256], "length": lc[:256]})])
        assert model.test_reversibility(xc[:64], ac[:64], lc[:64])
        u = torch.rand(B * N, 1, 2, generator=gen).cuda()
        z, ldj = model(xc, adjacency=ac, length=lc, u_noise=u)
        assert torch.isfinite(z).all() and torch.isfinite(ldj).all()
        h = B // 2
        z0, l0 = model(xc[:h], adjacency=ac[:h], length=lc[:h], u_noise=u[:h * N])
        assert_close(z0, z[:h], rtol=1e-4, atol=1e-4, what="batch split z")
        assert_close(l0, ldj[:h], rtol=1e-4, atol=1e-3, what="batch split ldj")
        pad = (torch.arange(N)[None, :] < length[:, None]).cuda()
        assert (z[~pad] == 0).all()


# ---- Edge-GNN (GraphCNF steps 2 and 3) ------------------------------------------------------------------------------
def _build_edge_gnn(g):
    from categoricalnf_b200.layers.networks import (Edge2NodeAttnLayer, Edge2NodeQKVAttnLayer, EdgeGNN, EdgeGNNLayer,
                                                    Node2EdgePlainLayer)
    e2n = (lambda: Edge2NodeQKVAttnLayer(hidden_size_nodes=g.hn, hidden_size_edges=g.he, skip_config=2)) if g.qkv else \
        (lambda: Edge2NodeAttnLayer(hidden_size_nodes=g.hn, hidden_size_edges=g.he, skip_config=2))
    n2e = lambda: Node2EdgePlainLayer(hidden_size_nodes=g.hn, hidden_size_edges=g.he, skip_config=2)
    net = EdgeGNN(g.c_in_nodes, g.c_in_edges, g.c_out_nodes, g.c_out_edges, lambda: EdgeGNNLayer(e2n, n2e), num_layers=g.layers,
                  max_neighbours=g.max_neighbours)
    net.load_state_dict(_sd(g), strict=True)
    return net.cuda().eval()


EDGE_CASES = ["edge_gnn_attn_sparse", "edge_gnn_attn_dense", "edge_gnn_qkv_dense", "edge_gnn_qkv_sparse"]


@pytest.mark.parametrize("name", EDGE_CASES)
def test_edge_gnn_golden(name):
    g = load_golden(name)
    net = _build_edge_gnn(g)
    binary = (g.adjacency > 0).long().cuda() if g.sparse else None
    kw = dict(length=g.length.cuda(), x_indices=(g.x_indices1.cuda(), g.x_indices2.cuda()), mask_valid=g.mask_valid.cuda(),
              channel_padding_mask=g.pad.cuda(), binary_adjacency=binary)
    with torch.no_grad():
        nodes, edges = net(g.z_nodes.cuda(), g.z_edges.cuda(), **kw)
    assert_close(nodes, g.nodes_out, rtol=1e-4, atol=2e-5, what="nodes_out")
    assert_close(edges, g.edges_out, rtol=1e-4, atol=2e-5, what="edges_out")
    # reference calling convention of a single layer: full [B,P,He] edge tensor in and out
    layer = net.layers[0]
    gen = torch.Generator().manual_seed(3)
    nf = torch.randn(g.z_nodes.shape[0], g.z_nodes.shape[1], g.hn, generator=gen).cuda()
    ef = (torch.randn(g.z_edges.shape[0], g.z_edges.shape[1], g.he, generator=gen) * g.mask_valid.unsqueeze(-1)).cuda()
    with torch.no_grad():
        n1, e1 = layer(nf, ef, kw["x_indices"], kw["mask_valid"])
    assert e1.shape == ef.shape and bool((e1[kw["mask_valid"] == 0] == 0).all())
    # training path (dense differentiable glue) agrees with the kernels
    with torch.enable_grad():
        nodes_t, edges_t = net(g.z_nodes.cuda().requires_grad_(True), g.z_edges.cuda(), **kw)
    assert_close(nodes_t, nodes, rtol=1e-4, atol=2e-5, what="training path nodes")
    assert_close(edges_t, edges, rtol=1e-4, atol=2e-5, what="training path edges")


@pytest.mark.parametrize("qkv", [False, True])
def test_edge_gnn_zinc_shape_vs_oracle(qkv):
    """GraphCNF hyper-parameters at the per-GPU molecule shape: B 64, N 38 (703 pairs), hidden 384 / 192, 4 layers."""
    from categoricalnf_b200.layers.networks import (Edge2NodeAttnLayer, Edge2NodeQKVAttnLayer, EdgeGNN, EdgeGNNLayer,
                                                    Node2EdgePlainLayer)
    torch.manual_seed(int(qkv))
    gen = torch.Generator().manual_seed(10 + int(qkv))
    B, N, hn, he = 64, 38, 384, 192
    e2n = (lambda: Edge2NodeQKVAttnLayer(hn, he, skip_config=2)) if qkv else (lambda: Edge2NodeAttnLayer(hn, he, skip_config=2))
    net = EdgeGNN(6, 2, 6 * 50, 2 * 26, lambda: EdgeGNNLayer(e2n, lambda: Node2EdgePlainLayer(hn, he, skip_config=2)),
                  num_layers=4, max_neighbours=4).eval()
    adj, length = _graphs(gen, B, N, 3, p=0.08)
    i1 = torch.tensor([i for i in range(N) for j in range(i + 1, N)])
    i2 = torch.tensor([j for i in range(N) for j in range(i + 1, N)])
    mask_all = ((i1[None, :] < length[:, None]) & (i2[None, :] < length[:, None])).float()
    bonds = adj.view(B, N * N).index_select(1, i1 + i2 * N) != 0
    mask_valid = mask_all if qkv else mask_all * bonds.float()
    pad = (torch.arange(N)[None, :] < length[:, None]).float().unsqueeze(-1)
    z_nodes = torch.randn(B, N, 6, generator=gen) * pad
    z_edges = torch.randn(B, i1.numel(), 2, generator=gen) * mask_valid.unsqueeze(-1)
    binary = None if qkv else (adj > 0).long()
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    ref_n, ref_e = GO.edge_gnn(sd, z_nodes, z_edges, (i1, i2), mask_valid, num_layers=4, qkv=qkv, pad=pad, binary_adjacency=binary,
                               max_neighbours=4)
    net = net.cuda()
    with torch.no_grad():
        nodes, edges = net(z_nodes.cuda(), z_edges.cuda(), length=length.cuda(), x_indices=(i1.cuda(), i2.cuda()),
                           mask_valid=mask_valid.cuda(), channel_padding_mask=pad.cuda(),
                           binary_adjacency=None if binary is None else binary.cuda())
    assert_close(nodes, ref_n, rtol=1e-4, atol=2e-5, what="nodes_out")
    assert_close(edges, ref_e, rtol=1e-4, atol=2e-5, what="edges_out")


# ---- GraphCNF (BASELINE configs 4 and 5) ----------------------------------------------------------------------------
def _build_graphcnf(N, node_types=5, hidden=(32, 16), flows="1,2,2", layers=2, mixtures=(8, 4), sd=None):
    import numpy as np
    from categoricalnf_b200.experiments.molecule_generation import GraphCNF
    node_prior = np.log(np.array([0.5, 0.2, 0.15, 0.1, 0.05], dtype=np.float32)) if node_types == 5 else \
        np.zeros(node_types, dtype=np.float32)

    class _Dataset:
        max_num_nodes = staticmethod(lambda: N)
        num_node_types = staticmethod(lambda: node_types)
        num_edge_types = staticmethod(lambda: 3)
        num_max_neighbours = staticmethod(lambda: 4)
        get_node_prior = staticmethod(lambda data_root="data/": node_prior)
        get_edge_prior = staticmethod(lambda data_root="data/": np.log(np.array([0.7, 0.2, 0.1], dtype=np.float32)))

    enc = lambda d: {"use_dequantization": False, "use_variational": False, "use_decoder": False, "num_dimensions": d,
                     "flow_config": {"num_flows": 0, "hidden_layers": 2, "hidden_size": 128},
                     "decoder_config": {"num_layers": 1, "hidden_size": 64}}
    params = {"categ_encoding_nodes": enc(6), "categ_encoding_edges": enc(2), "coupling_hidden_size_nodes": hidden[0],
              "coupling_hidden_size_edges": hidden[1], "coupling_num_flows": flows, "coupling_hidden_layers": layers,
              "coupling_num_mixtures_nodes": mixtures[0], "coupling_num_mixtures_edges": mixtures[1], "coupling_mask_ratio": 0.5,
              "coupling_dropout": 0.0, "encoding_virtual_num_flows": 0}
    model = GraphCNF(params, _Dataset)
    if sd is not None:
        model.load_state_dict(sd, strict=True)
    return model.cuda().eval()


def test_graphcnf_golden():
    """Forward (log-likelihood incl. the prior of the edge latents) and reverse (sampling) of the reference's GraphCNF."""
    g = load_golden("graphcnf_small")
    model = _build_graphcnf(g.N, sd=_sd(g))
    with torch.no_grad():
        z, ldj = model(g.x.cuda(), adjacency=g.adjacency.cuda(), length=g.length.cuda(), u_noise=g.u_nodes.cuda(),
                       u_noise_edges=g.u_edges.cuda(), u_noise_virtual=g.u_virtual.cuda())
        assert_close(z, g.z, what="z nodes")
        assert_close(ldj, g.ldj, rtol=1e-4, atol=5e-4, what="ldj")
        (x_smp, adj_smp), ldj_smp = model(g.z_nodes_init.cuda(), reverse=True, length=g.length.cuda(), z_edges_init=g.z_edges_init.cuda())
    valid = torch.arange(g.N)[None, :] < g.length[:, None]
    assert torch.equal(adj_smp.cpu(), g.adj_smp)
    assert torch.equal(x_smp.cpu()[valid], g.x_smp[valid])
    assert_close(ldj_smp, g.ldj_smp, rtol=1e-4, atol=5e-4, what="ldj of the sampling pass")


def test_graphcnf_zinc_shape_properties():
    """BASELINE config 4 per-GPU shape (B 64, N 38, 9 node types, hidden 384 / 192, flows 4,6,6, 4 layers, K 16 / 8):
    data-dependent init, forward, batch-split consistency, sampling produces symmetric adjacencies on valid nodes only."""
    torch.manual_seed(0)
    gen = torch.Generator().manual_seed(0)
    B, N = 64, 38
    model = _build_graphcnf(N, node_types=9, hidden=(384, 192), flows="4,6,6", layers=4, mixtures=(16, 8))
    adj, length = _graphs(gen, B, N, 3, p=0.06)
    x = torch.randint(0, 9, (B, N), generator=gen) * (torch.arange(N)[None, :] < length[:, None]).long()
    xc, ac, lc = x.cuda(), adj.cuda(), length.cuda()
    P = N * (N - 1) // 2
    with torch.no_grad():
        model.initialize_data_dependent([(xc[:32], {"adjacency": ac[:32], "length": lc[:32]})])
        noise = dict(u_noise=torch.rand(B * N, 1, 6, generator=gen).cuda(), u_noise_edges=torch.rand(B * P, 1, 2, generator=gen).cuda(),
                     u_noise_virtual=torch.rand(B * P, 1, 2, generator=gen).cuda())
        z, ldj = model(xc, adjacency=ac, length=lc, **noise)
        assert torch.isfinite(z).all() and torch.isfinite(ldj).all()
        h = B // 2
        half = dict(u_noise=noise["u_noise"][:h * N], u_noise_edges=noise["u_noise_edges"][:h * P],
                    u_noise_virtual=noise["u_noise_virtual"][:h * P])
        z0, l0 = model(xc[:h], adjacency=ac[:h], length=lc[:h], **half)
        assert_close(z0, z[:h], rtol=1e-4, atol=1e-4, what="batch split z")
        assert_close(l0, ldj[:h], rtol=1e-4, atol=5e-3, what="batch split ldj")
        z_nodes = model.prior_distribution.sample(shape=(B, N, 6)) * (torch.arange(N, device="cuda")[None, :, None] < lc[:, None, None])
        (x_smp, adj_smp), ldj_smp = model(z_nodes, reverse=True, length=lc)
        assert x_smp.shape == (B, N) and adj_smp.shape == (B, N, N)
        assert torch.equal(adj_smp, adj_smp.transpose(1, 2)) and int(adj_smp.max()) <= 3
        valid = (torch.arange(N, device="cuda")[None, :] < lc[:, None])
        assert bool((adj_smp * (~(valid[:, :, None] & valid[:, None, :])).long() == 0).all())
        assert torch.isfinite(ldj_smp).all()


def test_graphcnf_cuda_graph_replay_matches_eager():
    """GraphedLogLikelihood: the forward pass captured once per (batch shape, padded pair-row counts) and replayed - same
    z and log-likelihood as the eager pass on the same noise, for the batch it was captured on, for other batches of the
    same bucket (replay only) and for a batch that needs a new capture; golden parity through the replay as well."""
    from categoricalnf_b200.experiments.molecule_generation import GraphedLogLikelihood
    g = load_golden("graphcnf_small")
    model = _build_graphcnf(g.N, sd=_sd(g))
    graphed = GraphedLogLikelihood(model, bucket=16)
    x, adj, length = g.x.cuda(), g.adjacency.cuda(), g.length.cuda()
    noise = dict(u_noise=g.u_nodes.cuda(), u_noise_edges=g.u_edges.cuda(), u_noise_virtual=g.u_virtual.cuda())
    with torch.no_grad():
        z, ldj = graphed(x, adj, length, **noise)
        assert_close(z, g.z, what="z nodes (graph replay vs reference golden)")
        assert_close(ldj, g.ldj, rtol=1e-4, atol=5e-4, what="ldj (graph replay vs reference golden)")
        gen = torch.Generator().manual_seed(5)
        for trial in range(4):
            adj2, len2 = _graphs(gen, x.shape[0], g.N, 3, p=0.25 + 0.1 * (trial % 2))
            x2 = torch.randint(0, 5, x.shape, generator=gen) * (torch.arange(g.N)[None, :] < len2[:, None]).long()
            P = g.N * (g.N - 1) // 2
            nz = dict(u_noise=torch.rand(x.shape[0], g.N, 6, generator=gen).cuda(), u_noise_edges=torch.rand(x.shape[0], P, 2, generator=gen).cuda(),
                      u_noise_virtual=torch.rand(x.shape[0], P, 2, generator=gen).cuda())
            z_e, ldj_e = model(x2.cuda(), adjacency=adj2.cuda(), length=len2.cuda(), **nz)
            z_g, ldj_g = graphed(x2.cuda(), adj2.cuda(), len2.cuda(), **nz)
            assert_close(z_g, z_e, rtol=1e-5, atol=1e-6, what="z (trial %d)" % trial)
            assert_close(ldj_g, ldj_e, rtol=1e-5, atol=1e-4, what="ldj (trial %d)" % trial)
        z_r, ldj_r = graphed(x, adj, length)          # internal noise: finite, different draw
        assert torch.isfinite(ldj_r).all() and not torch.equal(ldj_r, ldj)
    assert 1 <= graphed.captures <= 5


# ---- BASELINE configs 3 / 4 at FULL size: random samples against the oracle / the unmodified reference ---------------------
def test_graph_node_flow_full_size_spot_check_vs_oracle():
    """Config 3 (B 1024, N 20, 8 flows, hidden 384, 4 layers, K 8): the model runs the whole batch, 8 random graphs are
    recomputed by the oracle on the same noise (graphs are independent samples)."""
    import graph_workloads as G
    dev_ = torch.device("cuda", 0)
    model = G.build_gc_model(dev_, seed=0)
    x, adj, length = G.gc_graphs(torch.Generator().manual_seed(5), G.GC["B"])
    x0, a0, l0 = G.gc_graphs(torch.Generator().manual_seed(7), G.GC["init_batch"])
    G.data_init(model, x0.cuda(), a0.cuda(), l0.cuda())
    B, N = x.shape
    u = torch.rand(B * N, 1, 2, generator=torch.Generator().manual_seed(6))
    with torch.no_grad():
        z, ldj = model(x.cuda(), adjacency=adj.cuda(), length=length.cuda(), u_noise=u.cuda())
    idx = torch.randperm(B, generator=torch.Generator().manual_seed(8))[:8].sort().values
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    u_s = u.view(B, N, 1, 2)[idx].reshape(-1, 1, 2)
    z_ref, ldj_ref = GO.graph_node_flow(sd, x[idx], adj[idx], length[idx], u_s, num_flows=G.GC["flows"], num_layers=G.GC["layers"],
                                        num_mixtures=G.GC["K"])
    assert_close(z[idx.cuda()], z_ref, rtol=1e-4, atol=2e-5, what="z (full size, 8 graphs)")
    assert_close(ldj[idx.cuda()], ldj_ref, rtol=1e-4, atol=2e-4, what="ldj")


@pytest.mark.skipif(not __import__("workload").reference_available(), reason="baseline/_ref absent (tools/vendor_reference.sh)")
def test_graphcnf_full_size_spot_check_vs_reference():
    """Config 4 (GraphCNF at the Zinc shape, batch 512 on one GPU): 4 random molecules recomputed by the UNMODIFIED reference
    model (baseline/_ref, state dict loaded by name) on the same noise; and config 5's reverse pass on 2 sets of latents."""
    import graph_workloads as G
    dev_ = torch.device("cuda", 0)
    N = G.MOL["N"]
    P = N * (N - 1) // 2
    model = G.build_mol_model(dev_, seed=0)
    x0, a0, l0 = G.molecules(torch.Generator().manual_seed(7), G.MOL["init_batch"])
    G.data_init(model, x0.cuda(), a0.cuda(), l0.cuda())
    B = 512
    x, adj, length = G.molecules(torch.Generator().manual_seed(11), B)
    gen = torch.Generator().manual_seed(12)
    un, ue, uv = torch.rand(B * N, 1, 6, generator=gen), torch.rand(B * P, 1, 2, generator=gen), torch.rand(B * P, 1, 2, generator=gen)
    with torch.no_grad():
        z, ldj = model(x.cuda(), adjacency=adj.cuda(), length=length.cuda(), u_noise=un.cuda(), u_noise_edges=ue.cuda(),
                       u_noise_virtual=uv.cuda())
    idx = torch.randperm(B, generator=torch.Generator().manual_seed(13))[:4].sort().values
    ref = G.build_reference_like(model, "mol")
    draws = iter([un.view(B, N, 1, 6)[idx].reshape(-1, 1, 6), ue.view(B, P, 1, 2)[idx].reshape(-1, 1, 2),
                  uv.view(B, P, 1, 2)[idx].reshape(-1, 1, 2)])
    for e in G.reference_encodings(ref, "mol"):
        e.prior_distribution.distribution.sample = lambda sample_shape=torch.Size(): next(draws)
    with torch.no_grad():
        z_ref, ldj_ref = ref(x[idx], adjacency=adj[idx], length=length[idx])
    assert_close(z[idx.cuda()], z_ref, rtol=1e-4, atol=4e-5, what="z nodes (full size, 4 molecules)")
    assert_close(ldj[idx.cuda()], ldj_ref, rtol=1e-4, atol=5e-4, what="ldj")
    rel = ((ldj[idx.cuda()].cpu().double() - ldj_ref.double()).abs() / ldj_ref.double().abs()).max().item()
    assert rel <= 1e-4, "pure relative ldj deviation %.3e" % rel


def test_coupling_networks_run_once_per_forward():
    """The evaluation-time fast paths of FlowModel (accumulate / next-block forms) must decide BEFORE a coupling network
    has been evaluated whether they apply: a fallback after the fact would run the RGCN twice (caught by the N = 2 bench
    run of round 2: graph colouring 18 -> 36 ms).  ``RGCNNet.cnf_features`` is the body of the network on every path
    (``forward`` calls it as well), so it must run exactly once per coupling layer and pass."""
    from categoricalnf_b200.layers.networks import graph_layers as GL
    g = load_golden("graph_node_flow")
    model = _build_flow(g)
    orig = GL.RGCNNet.cnf_features
    calls = {"n": 0}

    def counted(self, *a, **k):
        calls["n"] += 1
        return orig(self, *a, **k)

    GL.RGCNNet.cnf_features = counted
    try:
        with torch.no_grad():
            z, ldj = model(g.x.cuda(), adjacency=g.adjacency.cuda(), length=g.length.cuda(), u_noise=g.u.cuda())
    finally:
        GL.RGCNNet.cnf_features = orig
    n_couplings = sum(1 for layer in model.flow_layers if isinstance(getattr(layer, "nn", None), GL.RGCNNet))
    assert n_couplings >= 2 and calls["n"] == n_couplings, "RGCN bodies evaluated %d times for %d couplings" % (calls["n"], n_couplings)
    assert_close(ldj, g.ldj, rtol=1e-4, atol=2e-4, what="ldj")
