"""CPU oracle of the node-based graph coupling networks (TEST INFRASTRUCTURE ONLY - imported by tests/,
__graft_entry__.smoke() and bench.py's CPU legs; nothing under categoricalnf_b200/ may import it).

Restates, as plain dense tensor algebra on state-dict tensors, what the reference computes with one-hot adjacencies
and padded neighbour gathers:
    RelationGraphConv       layers/networks/graph_layers.py:15-50
    RelationGraphAttention  layers/networks/graph_layers.py:53-154
    GNNSkipConnection       layers/networks/graph_layers.py:702-733
    RGCNNet                 layers/networks/graph_layers.py:157-235
    GraphNodeFlow.forward   experiments/graph_coloring/graph_node_flow.py:17-102 (flow container: flow_model.py:25-53)
Pinned by tests/test_oracle_golden.py against tests/golden/rgcn_*.npz and graph_node_flow.npz, which were produced by
the unmodified reference classes (tests/golden/make_golden.py).  float32 like the reference networks.
"""
import torch
import torch.nn.functional as F

from . import cnf_oracle as O


def _lin(sd, key, x):
    return F.linear(x, sd[key + ".weight"], sd[key + ".bias"])


def _ln(sd, key, x):
    return F.layer_norm(x, (x.shape[-1],), sd[key + ".weight"], sd[key + ".bias"], 1e-5)


def relation_graph_conv(sd, pre, x, adj, num_edges, num_neighbours):
    """graph_layers.py:37-50.  ``adj`` integer [B,N,N]; the neighbour sum runs over the FIRST node index (:45)."""
    B, N, _ = x.shape
    xn = _ln(sd, pre + "norm_layer", x)
    hs = _lin(sd, pre + "linear_hs", xn)
    hr_all = _lin(sd, pre + "linear_hr", xn).view(B, N, num_edges, -1)
    out = hs.clone()
    denom = num_neighbours.clamp(min=1e-5)
    for e in range(num_edges):
        a = (adj == e + 1).float()                                   # [B,j,i]
        out = out + torch.einsum("bji,bjc->bic", a, hr_all[:, :, e]) / denom.unsqueeze(-1)
    return out


def relation_graph_attention(sd, pre, x, adj, num_edges, num_heads=4, slope=0.2):
    """graph_layers.py:76-154 as a dense masked softmax over all node pairs; the self-connection is edge slot E (:99-102)."""
    B, N, _ = x.shape
    xn = _ln(sd, pre + "norm_layer", x)
    aw = sd[pre + "attn_weight"]
    H, _, Dh = aw.shape
    assert H == num_heads
    hs = _lin(sd, pre + "linear_hs", xn).reshape(B, N, H, Dh)
    hr = _lin(sd, pre + "linear_hr", xn).reshape(B, N, num_edges + 1, H, Dh)
    hs_attn = (hs * aw[:, 0]).sum(-1)                                 # [B,i,H]
    hr_attn = (hr * aw[:, 1]).sum(-1)                                 # [B,j,E+1,H]
    eye = torch.eye(N, dtype=torch.bool).unsqueeze(0)
    etype = torch.where(eye, torch.full_like(adj, num_edges + 1), adj)   # [B,i,j], 0 = absent
    logits = torch.full((B, N, N, H), float("-inf"))
    values = torch.zeros(B, N, N, H, Dh)
    for e in range(num_edges + 1):
        sel = (etype == e + 1)                                        # [B,i,j]
        l = F.leaky_relu(hs_attn[:, :, None, :] + hr_attn[:, None, :, e, :], slope)
        logits = torch.where(sel.unsqueeze(-1), l, logits)
        values = torch.where(sel[..., None, None], hr[:, None, :, e].expand(B, N, N, H, Dh), values)
    probs = torch.softmax(logits, dim=2)
    att = (probs.unsqueeze(-1) * values).sum(dim=2).reshape(B, N, H * Dh)
    return _lin(sd, pre + "output_projection.1", F.gelu(att))         # :68-71


def skip_connection(sd, pre, orig, feat, config):
    """graph_layers.py:722-733."""
    s = _lin(sd, pre + "skip_layer", feat)
    if config == 0:
        return orig + s
    val, gate = s.chunk(2, dim=-1)
    gate = torch.sigmoid(gate)
    return orig + val * gate if config == 1 else orig * (1 - gate) + val * gate


def rgcn_net(sd, x, adj, *, num_edges, num_layers, attention, skip_config=2, max_neighbours=4, pad=None, pre=""):
    """RGCNNet.forward (graph_layers.py:204-235)."""
    num_neighbours = (adj > 0).sum(dim=1).float()                     # one_hot(...)[...,1:].sum(dim=[1,3]) (:205-206)
    h = _lin(sd, pre + "input_layer.2", F.gelu(_lin(sd, pre + "input_layer.0", x)))
    if max_neighbours > 0:
        num_neighbours = num_neighbours.clamp(max=max_neighbours)     # :210 - the clamped count also feeds the layers
        onehot = F.one_hot(num_neighbours.long(), max_neighbours + 1).float()
        h = torch.cat([h, _lin(sd, pre + "neighbour_embed", onehot)], dim=-1)
    for i in range(num_layers):
        lp = "%slayers.%d." % (pre, i)
        if attention:
            f = relation_graph_attention(sd, lp + "0.", h, adj, num_edges)
        else:
            f = relation_graph_conv(sd, lp + "0.", h, adj, num_edges, num_neighbours)
        h = skip_connection(sd, lp + "3.", h, F.gelu(f), skip_config)
    h = _ln(sd, pre + "output_layer.0", h)
    out = _lin(sd, pre + "output_layer.3", F.gelu(_lin(sd, pre + "output_layer.1", h)))
    return out * pad if pad is not None else out


def graph_node_flow(sd, x, adj, length, u_noise, *, num_flows, num_layers, num_mixtures, num_node_types=3, reverse_z=None):
    """GraphNodeFlow.forward (graph_node_flow.py:98-102): encoding, ``num_flows`` x [ActNorm, InvConv,
    MixtureCDFCoupling(RGCNNet attention, num_edges 1)], final ActNorm; ldj accumulated as FlowModel does.
    ``reverse_z`` given -> instead run the continuous layers backwards on it and return (z, ldj)."""
    B, N = x.shape
    pad = (torch.arange(N)[None, :] < length[:, None]).float().unsqueeze(-1)
    D = sd["flow_layers.1.bias"].shape[-1]
    mask = O.expand_mask(O.channel_mask(D, 0.5), torch.zeros(B, N, D))

    def block(i):
        base = 1 + 3 * i
        w, sldj = O.invconv_weight(*(sd["flow_layers.%d.%s" % (base + 1, k)] for k in ("p", "l", "log_s", "u", "sign_s")))
        return base, w, sldj

    def coupling(idx, z, reverse):
        pre = "flow_layers.%d." % idx
        nn_out = rgcn_net(sd, z * mask, adj, num_edges=1, num_layers=num_layers, attention=True, pre=pre + "nn.")
        return O.mixcdf_coupling(z, nn_out, mask, num_mixtures, sd[pre + "scaling_factor"], sd[pre + "mixture_scaling_factor"],
                                 reverse=reverse, pad=pad, reg_max=3.5, reg_factor=2.0, training=False)

    last = 1 + 3 * num_flows
    if reverse_z is None:
        table = O.categ_table(sd["flow_layers.0.embed_layer.weight"], sd["flow_layers.0.flow_layers.0.pred_net.layer.weight"],
                              sd["flow_layers.0.flow_layers.0.pred_net.layer.bias"])
        z, ldj, _ = O.categ_encode(x, u_noise, table, sd["flow_layers.0.category_prior"], pad=pad)
        for i in range(num_flows):
            base, w, sldj = block(i)
            z, ldj = O.actnorm(z, sd["flow_layers.%d.bias" % base], sd["flow_layers.%d.scales" % base], ldj, length=length, pad=pad)
            z, ldj = O.invconv(z, w, sldj, ldj, length=length, pad=pad)
            z, l, _ = coupling(base + 2, z, False)
            ldj = ldj + l
        z, ldj = O.actnorm(z, sd["flow_layers.%d.bias" % last], sd["flow_layers.%d.scales" % last], ldj, length=length, pad=pad)
        return z, ldj
    z, ldj = reverse_z, torch.zeros(B)
    z, ldj = O.actnorm(z, sd["flow_layers.%d.bias" % last], sd["flow_layers.%d.scales" % last], ldj, reverse=True, length=length, pad=pad)
    for i in reversed(range(num_flows)):
        base, w, sldj = block(i)
        z, l, _ = coupling(base + 2, z, True)
        ldj = ldj + l
        z, ldj = O.invconv(z, O.invconv_inverse(w), sldj, ldj, reverse=True, length=length, pad=pad)
        z, ldj = O.actnorm(z, sd["flow_layers.%d.bias" % base], sd["flow_layers.%d.scales" % base], ldj, reverse=True, length=length, pad=pad)
    return z, ldj


# ---------------------------------------------------------------------------------------------------------------------
# Edge-GNN (layers/networks/graph_layers.py:242-820) as dense [B,N,N] algebra.  Pinned by tests/golden/edge_gnn_*.npz
# (reference EdgeGNN, dense and sparse forward passes).
# ---------------------------------------------------------------------------------------------------------------------
def _to_matrix(pair_vals, x_indices, N):
    """[B,P,*] pair list -> symmetric [B,N,N,*] (zero diagonal)."""
    B = pair_vals.shape[0]
    out = pair_vals.new_zeros((B, N, N) + pair_vals.shape[2:])
    out[:, x_indices[0], x_indices[1]] = pair_vals
    out[:, x_indices[1], x_indices[0]] = pair_vals
    return out


def _mlp_in(sd, pre, x):
    return _lin(sd, pre + ".2", F.gelu(_lin(sd, pre + ".0", x)))


def _mlp_out(sd, pre, x):
    return _lin(sd, pre + ".3", F.gelu(_lin(sd, pre + ".1", _ln(sd, pre + ".0", x))))


def edge2node_attn(sd, pre, node, edge, x_indices, mask_valid, H=4):
    """Edge2NodeAttnLayer (:595-645)."""
    B, N, _ = node.shape
    new = _lin(sd, pre + "node_feat_layer", _ln(sd, pre + "node_normalization", node))
    node_self, node_ctx = new.chunk(2, dim=-1)
    edge_in = _ln(sd, pre + "edge_normalization", edge)
    e_new = _to_matrix(_lin(sd, pre + "edge_feat_layer", edge_in), x_indices, N)          # [B,i,j,HO]
    e_log = _to_matrix(_lin(sd, pre + "edge_logits_layer", edge_in), x_indices, N)        # [B,i,j,H]
    m = _to_matrix(mask_valid, x_indices, N).unsqueeze(-1)
    w = torch.sigmoid(e_log) * m
    probs = w / w.sum(dim=2, keepdim=True).clamp(min=1e-5)                                # :628-629
    Dh = node_ctx.shape[-1] // H
    pair_feat = (e_new + node_ctx[:, None, :, :]).view(B, N, N, H, Dh)
    att = (pair_feat * probs.unsqueeze(-1)).sum(dim=2).reshape(B, N, H * Dh)
    return skip_connection(sd, pre + "skip_layer.", node, F.gelu(node_self + att), 2)


def edge2node_qkv(sd, pre, node, edge, x_indices, mask_valid, H=4):
    """Edge2NodeQKVAttnLayer (:432-502)."""
    B, N, Hn = node.shape
    Dh = Hn // H
    node_in = _ln(sd, pre + "node_normalization", node)
    edge_in = _ln(sd, pre + "edge_normalization", edge)
    q, k, v = (t.view(B, N, H, Dh) for t in _lin(sd, pre + "node_query_key_val_layer", node_in).chunk(3, dim=-1))
    e_val = _to_matrix(_lin(sd, pre + "edge_val_layer", edge_in), x_indices, N).view(B, N, N, H, Dh)
    e_adj = _to_matrix(_lin(sd, pre + "edge_adj_layer", edge_in), x_indices, N)           # [B,i,j,H]
    m = _to_matrix(mask_valid, x_indices, N).unsqueeze(-1)
    logits = torch.einsum("bihd,bjhd->bijh", q, k) * float(Dh) ** -0.5 + e_adj            # :446,:476
    logits = logits.masked_fill(m == 0, -9e15)
    probs = torch.softmax(logits, dim=2) * (m.sum(dim=2, keepdim=True) > 0).float()       # :477-478
    att = ((e_val + v[:, None]) * probs.unsqueeze(-1)).sum(dim=2).reshape(B, N, H * Dh)
    comb = F.gelu(_lin(sd, pre + "output_projection", torch.cat([node_in, att], dim=-1)))
    return skip_connection(sd, pre + "skip_layer.", node, comb, 2)


def node2edge(sd, pre, node, edge, x_indices, mask_valid):
    """Node2EdgePlainLayer (:317-336) + the masking of EdgeGNNLayer (:261)."""
    nl = _lin(sd, pre + "node_feat_layer.1", _ln(sd, pre + "node_feat_layer.0", node))
    el = _lin(sd, pre + "edge_feat_layer.1", _ln(sd, pre + "edge_feat_layer.0", edge))
    comb = F.gelu(el + nl[:, x_indices[0]] + nl[:, x_indices[1]])
    return skip_connection(sd, pre + "skip_layer.", edge, comb, 2) * mask_valid.unsqueeze(-1)


def edge_gnn(sd, z_nodes, z_edges, x_indices, mask_valid, *, num_layers, qkv, pad=None, binary_adjacency=None, max_neighbours=-1, pre=""):
    """EdgeGNN.forward (:781-820) -> (nodes_out, edges_out)."""
    node = _mlp_in(sd, pre + "input_layer_nodes", z_nodes)
    edge = _mlp_in(sd, pre + "input_layer_edges", z_edges)
    if binary_adjacency is not None and max_neighbours > 0:
        nn_ = binary_adjacency.sum(dim=-1).long().clamp(max=max_neighbours)
        node = node + _lin(sd, pre + "node_neighbour_embed", F.one_hot(nn_, max_neighbours + 1).float())
    for i in range(num_layers):
        lp = "%slayers.%d." % (pre, i)
        node = (edge2node_qkv if qkv else edge2node_attn)(sd, lp + "edge2node_layer.", node, edge, x_indices, mask_valid)
        edge = node2edge(sd, lp + "node2edge_layer.", node, edge, x_indices, mask_valid)
    nodes_out = _mlp_out(sd, pre + "out_layer_nodes", node)
    edges_out = _mlp_out(sd, pre + "out_layer_edges", edge) * mask_valid.unsqueeze(-1)
    return (nodes_out * pad if pad is not None else nodes_out), edges_out
