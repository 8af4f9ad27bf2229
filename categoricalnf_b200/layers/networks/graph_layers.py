"""Node-based graph coupling networks (reference layers/networks/graph_layers.py:15-235, 702-733).

``RGCNNet`` is the coupling network of GraphCNF's node flows: graph colouring uses it with
``RelationGraphAttention`` layers (experiments/graph_coloring/graph_node_flow.py:53-59), molecule generation
step 1 with ``RelationGraphConv`` (experiments/molecule_generation/graphCNF.py:104-111).  Same constructors,
attribute and parameter names as the reference, so its checkpoints load (``layers.<i>.0.linear_hs.weight``,
``layers.<i>.3.skip_layer.weight``, ``input_layer.0.weight``, ``output_layer.3.weight`` ...).

What runs where (evaluation / sampling, i.e. grad disabled):
  * every projection is a ``TCLinear`` (an ``nn.Linear`` subclass, same state-dict keys) -> ``cnf_linear_fwd`` (tcgen05, 3xTF32; hs and hr projections of a layer as ONE GEMM over
    the concatenated weights; GELU fused into the epilogue where the reference applies it next);
  * ``nn.LayerNorm``     -> ``cnf_layernorm``;
  * neighbour handling   -> ``cnf_graph_attn_scores`` + ``cnf_graph_aggregate`` working on the integer adjacency
    (the reference one-hot encodes it, pads every node to the batch-wide maximum degree with masked_select /
    index_select - one host sync per layer, :107 - and materialises the padded products);
  * ``GNNSkipConnection``-> ``cnf_skip_gate``.
With grad enabled (training) the projections still run on the tensor cores (``TCLinear`` forward + ``cnf_linear_bwd``)
and the glue is evaluated with differentiable dense torch operations of the same mathematics.
The *final* projection is exposed through ``cnf_features`` / ``cnf_final_linear`` so that ``MixtureCDFCoupling`` fuses
it with the transform (``cnf_linear_mixcdf_fwd``).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import ops
from .linear import TCLinear, _TCLinearFn

PRECISION = "3xtf32"


def _grad_mode(*tensors):
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


def _linear(x, lin, activation=None, precision=None):
    """``lin(x)`` (+ GELU) on the tensor cores; differentiable when grad is enabled."""
    precision = precision or PRECISION
    if _grad_mode(x, lin.weight, lin.bias):
        y = _TCLinearFn.apply(x, lin.weight, lin.bias, precision)
        return F.gelu(y) if activation == "gelu" else y
    return ops.linear(x, lin.weight, lin.bias, precision=precision, activation=activation)


def _layernorm(x, ln):
    if _grad_mode(x, ln.weight, ln.bias):
        return F.layer_norm(x, ln.normalized_shape, ln.weight, ln.bias, ln.eps)
    return ops.layernorm(x, ln.weight, ln.bias, ln.eps)


def _edge_types(adjacency, num_edges):
    """Integer edge types [B,N,N] (0 = no edge) from either the integer adjacency or the reference's one-hot form
    [B,N,N,E] (graph_layers.py:205)."""
    if adjacency.dim() == 4:
        idx = torch.arange(1, adjacency.shape[-1] + 1, device=adjacency.device, dtype=adjacency.dtype)
        return (adjacency * idx).sum(dim=-1).long()
    return adjacency.long()


class _FusedPair:
    """hs / hr projections of one layer as a single GEMM: concatenated weight and bias, rebuilt when a parameter changes."""

    def __init__(self):
        self.key, self.weight, self.bias = None, None, None

    def get(self, a, b):
        key = (a.weight.data_ptr(), a.weight._version, a.bias._version, b.weight.data_ptr(), b.weight._version,
               b.bias._version, a.weight.device)
        if key != self.key:
            with torch.no_grad():
                self.weight = torch.cat([a.weight, b.weight], dim=0).contiguous()
                self.bias = torch.cat([a.bias, b.bias], dim=0).contiguous()
            self.key = key
        return self.weight, self.bias


class RelationGraphConv(nn.Module):
    """h_i = W_s LN(x_i) + 1/n_i sum_j W_{e(j,i)} LN(x_j)   (graph_layers.py:15-50)."""

    def __init__(self, c_in, c_out, num_edges, **kwargs):
        super().__init__()
        self.c_in, self.c_out, self.num_edges = c_in, c_out, num_edges
        self.norm_layer = nn.LayerNorm(self.c_in)
        self.linear_hs = TCLinear(self.c_in, self.c_out)
        self.linear_hr = TCLinear(self.c_in, self.c_out * self.num_edges)
        self.__dict__["_pair"] = _FusedPair()

    def forward(self, x, adjacency, num_neighbours=None, activation=None, **kwargs):
        """``adjacency``: integer edge types [B,N,N] or one-hot [B,N,N,E]; ``num_neighbours`` [B,N] (None -> counted).
        ``activation="gelu"`` applies the GELU that follows the layer inside RGCNNet in the same kernel."""
        adj = _edge_types(adjacency, self.num_edges)
        if _grad_mode(x, self.linear_hs.weight):
            return self._forward_dense(x, adj, num_neighbours, activation)
        B, N = x.shape[0], x.shape[1]
        xn = ops.layernorm(x, self.norm_layer.weight, self.norm_layer.bias, self.norm_layer.eps)
        w, b = self._pair.get(self.linear_hs, self.linear_hr)
        y = ops.linear(xn.reshape(B * N, -1), w, b, precision=PRECISION)
        hs, hr = y[:, :self.c_out], y[:, self.c_out:]
        nn_ = None if num_neighbours is None else num_neighbours.reshape(B, N).float()
        return ops.graph_mean_aggregate(hs.unflatten(0, (B, N)), hr.unflatten(0, (B, N)), adj, self.num_edges, nn_,
                                        activation=activation)

    def _forward_dense(self, x, adj, num_neighbours, activation):
        B, N = x.shape[0], x.shape[1]
        xn = _layernorm(x, self.norm_layer)
        hs = _linear(xn, self.linear_hs)
        hr_all = _linear(xn, self.linear_hr).view(B, N, self.num_edges, self.c_out)
        onehot = F.one_hot(adj, self.num_edges + 1)[..., 1:].to(hs.dtype)                  # [B,j,i,E]
        if num_neighbours is None:
            num_neighbours = onehot.sum(dim=[1, 3])
        hr = torch.einsum("bjie,bjec->bic", onehot, hr_all) / num_neighbours.unsqueeze(-1).clamp(min=1e-5)
        out = hs + hr
        return F.gelu(out) if activation == "gelu" else out


class RelationGraphAttention(nn.Module):
    """Multi-head graph attention over typed edges plus a self-connection (graph_layers.py:53-154)."""

    def __init__(self, c_in, c_out, num_edges, num_heads=4, **kwargs):
        super().__init__()
        self.c_in, self.c_out, self.num_edges, self.num_heads = c_in, c_out, num_edges, num_heads
        self.c_out_per_head = self.c_out * 2 // self.num_heads
        width = self.c_out_per_head * self.num_heads
        self.norm_layer = nn.LayerNorm(self.c_in)
        self.linear_hs = TCLinear(self.c_in, width)
        self.linear_hr = TCLinear(self.c_in, width * (self.num_edges + 1))
        self.attn_weight = nn.Parameter(torch.zeros(self.num_heads, 2, self.c_out_per_head), requires_grad=True)
        nn.init.xavier_uniform_(self.attn_weight.data, gain=1.414)
        self.output_projection = nn.Sequential(nn.GELU(), TCLinear(width, self.c_out))
        self.leaky_relu = nn.LeakyReLU(0.2)
        self.__dict__["_pair"] = _FusedPair()

    def forward(self, x, adjacency, activation=None, **kwargs):
        adj = _edge_types(adjacency, self.num_edges)
        if _grad_mode(x, self.linear_hs.weight, self.attn_weight):
            return self._forward_dense(x, adj, activation)
        B, N = x.shape[0], x.shape[1]
        width = self.c_out_per_head * self.num_heads
        xn = ops.layernorm(x, self.norm_layer.weight, self.norm_layer.bias, self.norm_layer.eps)
        w, b = self._pair.get(self.linear_hs, self.linear_hr)
        y = ops.linear(xn.reshape(B * N, -1), w, b, precision=PRECISION)
        att = ops.graph_attention_aggregate(y[:, :width].unflatten(0, (B, N)), y[:, width:].unflatten(0, (B, N)),
                                            self.attn_weight, adj, self.num_edges,
                                            leaky_slope=self.leaky_relu.negative_slope, activation="gelu")
        return ops.linear(att, self.output_projection[1].weight, self.output_projection[1].bias, precision=PRECISION,
                          activation=activation)

    def _forward_dense(self, x, adj, activation):
        """Same mathematics as a dense masked softmax over all node pairs (differentiable; training path)."""
        B, N = x.shape[0], x.shape[1]
        H, Dh, E = self.num_heads, self.c_out_per_head, self.num_edges
        xn = _layernorm(x, self.norm_layer)
        hs = _linear(xn, self.linear_hs).reshape(B, N, H, Dh)
        hr_all = _linear(xn, self.linear_hr).reshape(B, N, E + 1, H, Dh)
        hs_attn = (hs * self.attn_weight[:, 0].view(1, 1, H, Dh)).sum(dim=-1)                       # [B,i,H]
        hr_attn = (hr_all * self.attn_weight[:, 1].view(1, 1, 1, H, Dh)).sum(dim=-1)                # [B,j,E+1,H]
        eye = torch.eye(N, device=x.device, dtype=torch.long).unsqueeze(0)
        etype = torch.where(eye.bool(), torch.full_like(adj, E + 1), adj)                           # self-connection = slot E
        present = etype > 0
        slot = (etype - 1).clamp(min=0)                                                             # [B,i,j]
        idx = slot.unsqueeze(-1).expand(B, N, N, H)
        hr_attn_ij = torch.gather(hr_attn.unsqueeze(1).expand(B, N, N, E + 1, H), 3, idx.unsqueeze(3)).squeeze(3)
        logits = self.leaky_relu(hs_attn.unsqueeze(2) + hr_attn_ij)
        logits = logits.masked_fill(~present.unsqueeze(-1), -9e15)
        probs = torch.softmax(logits, dim=2)                                                        # over j
        idx5 = slot.view(B, N, N, 1, 1, 1).expand(B, N, N, 1, H, Dh)
        hr_ij = torch.gather(hr_all.unsqueeze(1).expand(B, N, N, E + 1, H, Dh), 3, idx5).squeeze(3)  # [B,i,j,H,Dh]
        att = (probs.unsqueeze(-1) * hr_ij).sum(dim=2).reshape(B, N, H * Dh)
        out = _linear(F.gelu(att), self.output_projection[1])
        return F.gelu(out) if activation == "gelu" else out


class GNNSkipConnection(nn.Module):
    """Residual (0), gated (1) or highway (2) combination of a block's input and output (graph_layers.py:702-733)."""

    def __init__(self, hidden_size, config=0, input_size=-1, dp_rate=0.0):
        super().__init__()
        self.hidden_size = hidden_size
        self.input_size = input_size if input_size > 0 else hidden_size
        self.config = config
        self.dp_rate = dp_rate
        assert self.config in (0, 1, 2), "[!] ERROR: Unknown skip connection config \"%s\"" % str(self.config)
        self.skip_layer = TCLinear(self.input_size, self.hidden_size * (1 if self.config == 0 else 2))
        if self.dp_rate > 0.0:
            self.skip_layer = nn.Sequential(nn.Dropout(self.dp_rate), self.skip_layer)

    def _skip(self, feat):
        if isinstance(self.skip_layer, nn.Sequential):
            return _linear(self.skip_layer[0](feat), self.skip_layer[1])
        return _linear(feat, self.skip_layer)

    def forward(self, orig, feat):
        s = self._skip(feat)
        if not _grad_mode(orig, s):
            return ops.skip_gate(orig, s, self.config)
        if self.config == 0:
            return orig + s
        val, gate_logits = s.chunk(2, dim=-1)
        gate = torch.sigmoid(gate_logits)
        return orig + val * gate if self.config == 1 else orig * (1 - gate) + val * gate


class RGCNNet(nn.Module):

    def __init__(self, c_in, c_out, num_edges, num_layers, hidden_size, dp_rate=0.0, max_neighbours=4, skip_config=2,
                 rgc_layer_fun=RelationGraphConv, **kwargs):
        super().__init__()
        self.c_in, self.c_out, self.num_edges, self.num_layers = c_in, c_out, num_edges, num_layers
        self.hidden_size, self.dp_rate, self.max_neighbours = hidden_size, dp_rate, max_neighbours
        if self.max_neighbours > 0:
            neighbour_embed_size = int(hidden_size // 4)
            self.neighbour_embed = TCLinear(max_neighbours + 1, neighbour_embed_size)
        else:
            neighbour_embed_size = 0
        self.act_fn = nn.GELU()
        self.dropout = nn.Dropout(dp_rate)
        self.layers = nn.ModuleList([
            nn.ModuleList([rgc_layer_fun(c_in=hidden_size, c_out=hidden_size, num_edges=num_edges), self.act_fn, self.dropout,
                           GNNSkipConnection(hidden_size=hidden_size, config=skip_config)])
            for _ in range(num_layers)])
        self.input_layer = nn.Sequential(TCLinear(c_in, hidden_size), self.act_fn,
                                         TCLinear(hidden_size, hidden_size - neighbour_embed_size))
        self.output_layer = nn.Sequential(nn.LayerNorm(hidden_size), TCLinear(hidden_size, hidden_size), self.act_fn,
                                          TCLinear(hidden_size, c_out))

    # -- protocol of split_final_linear(): everything up to the last projection / that projection ------------------
    @property
    def cnf_final_linear(self):
        return self.output_layer[3]

    def cnf_features(self, x, adjacency=None, **kwargs):
        adj = _edge_types(adjacency, self.num_edges)
        # number of neighbours per node: edges counted over the FIRST node index, as upstream (:206)
        num_neighbours = (adj > 0).sum(dim=1).float()
        h = _linear(_linear(x, self.input_layer[0], activation="gelu"), self.input_layer[2])
        if self.max_neighbours > 0:
            num_neighbours = num_neighbours.clamp(max=self.max_neighbours)
            # Linear(one_hot(n)) = the n-th column of the weight + bias (:211)
            neigh = F.embedding(num_neighbours.long(), self.neighbour_embed.weight.t()) + self.neighbour_embed.bias
            h = torch.cat([h, neigh], dim=-1)
        for block in self.layers:
            rgc, skip = block[0], block[3]
            # block = [graph layer, GELU, dropout, skip]: the GELU is fused into the graph layer's last kernel
            feat = rgc(h, adjacency=adj, num_neighbours=num_neighbours, activation="gelu")
            feat = self.dropout(feat)
            h = skip(orig=h, feat=feat)
        h = _layernorm(h, self.output_layer[0])
        return _linear(h, self.output_layer[1], activation="gelu")

    def forward(self, x, adjacency, channel_padding_mask=None, embed_ext_input=None, **kwargs):
        out = _linear(self.cnf_features(x, adjacency=adjacency), self.output_layer[3])
        if channel_padding_mask is not None:
            out = out * channel_padding_mask
        return out
