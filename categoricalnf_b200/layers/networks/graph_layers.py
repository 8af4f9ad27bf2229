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
With grad enabled (training) the same kernels run inside ``autograd.Function``s (``graph_functional.py``) whose backward
passes are the kernels of ``csrc/graph_ops_bwd.cu`` (``cnf_layernorm_bwd``, ``cnf_graph_aggregate_bwd``, ``cnf_skip_gate_bwd``,
``cnf_edge_aggregate_bwd``, ``cnf_pair_combine_bwd``, ``cnf_gelu``) next to ``cnf_linear_bwd`` for the projections.
The *final* projection is exposed through ``cnf_features`` / ``cnf_final_linear`` so that ``MixtureCDFCoupling`` fuses
it with the transform (``cnf_linear_mixcdf_fwd``).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import graph_functional as GF
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
        return GF.gelu(y) if activation == "gelu" else y
    return ops.linear(x, lin.weight, lin.bias, precision=precision, activation=activation)


def _layernorm(x, ln):
    """nn.LayerNorm: ``cnf_layernorm`` forward, ``cnf_layernorm_bwd`` under autograd."""
    return GF.layernorm(x, ln.weight, ln.bias, ln.eps)


def _edge_types(adjacency, num_edges):
    """Integer edge types [B,N,N] (0 = no edge) from either the integer adjacency or the reference's one-hot form
    [B,N,N,E] (graph_layers.py:205)."""
    if adjacency.dim() == 4:
        idx = torch.arange(1, adjacency.shape[-1] + 1, device=adjacency.device, dtype=adjacency.dtype)
        return (adjacency * idx).sum(dim=-1).long()
    return adjacency.long()


class _FusedPair:
    """hs / hr projections of one layer as a single GEMM: concatenated weight and bias, rebuilt when a parameter changes.
    With ``attn_weight`` the attention logits ride along as extra output columns - they are linear in the projection
    input: hs_i . a_0 = x_i (W_hs^T a_0) + b_hs . a_0 (graph_layers.py:93-96) - so no separate pass reads hs / hr again.
    Column layout: [hs | hr | score_s (H) | score_r ((E+1) H) | zero padding to a multiple of 4]."""

    def __init__(self):
        self.key, self.weight, self.bias = None, None, None

    @staticmethod
    def build(a, b, attn_weight=None):
        """(weight, bias) of the fused projection as differentiable functions of the layer's parameters."""
        ws, bs = [a.weight, b.weight], [a.bias, b.bias]
        if attn_weight is not None:
            H, _, Dh = attn_weight.shape
            K = a.weight.shape[1]
            aw = attn_weight.double()
            ws.append(torch.einsum("hd,hdk->hk", aw[:, 0], a.weight.double().view(H, Dh, K)).float())
            bs.append((aw[:, 0] * a.bias.double().view(H, Dh)).sum(-1).float())
            E1 = b.weight.shape[0] // (H * Dh)
            ws.append(torch.einsum("hd,ehdk->ehk", aw[:, 1], b.weight.double().view(E1, H, Dh, K)).reshape(E1 * H, K).float())
            bs.append((aw[:, 1].unsqueeze(0) * b.bias.double().view(E1, H, Dh)).sum(-1).reshape(E1 * H).float())
            extra = (-(H + E1 * H)) % 4
            if extra:
                ws.append(a.weight.new_zeros(extra, K))
                bs.append(a.bias.new_zeros(extra))
        w = torch.cat(ws, dim=0).contiguous()
        w._cnf_cache_lo = True      # a weight: ops.weight_split splits it once per call (training) or once per version (get)
        return w, torch.cat(bs, dim=0).contiguous()

    def get(self, a, b, attn_weight=None):
        key = (a.weight.data_ptr(), a.weight._version, a.bias._version, b.weight.data_ptr(), b.weight._version,
               b.bias._version, a.weight.device, None if attn_weight is None else (attn_weight.data_ptr(), attn_weight._version),
               ops.param_epoch())
        if key != self.key:
            with torch.no_grad():
                self.weight, self.bias = self.build(a, b, attn_weight)
                self.weight._cnf_cache_lo = True       # long-lived: ops.linear may cache its TF32 low part
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
        B, N = x.shape[0], x.shape[1]
        if _grad_mode(x, self.linear_hs.weight):
            # training: same kernels, differentiable (cnf_layernorm_bwd, cnf_linear_bwd, cnf_graph_aggregate_bwd)
            xn = _layernorm(x, self.norm_layer)
            w, b = _FusedPair.build(self.linear_hs, self.linear_hr)
            y = _TCLinearFn.apply(xn.reshape(B * N, -1), w, b, PRECISION)
            cfg = dict(B=B, N=N, E=self.num_edges, H=1, Dh=self.c_out, mode=0, off_hs=0, off_hr=self.c_out, activation=activation)
            return GF.graph_aggregate(y, adj, cfg, num_neighbours)
        xn = ops.layernorm(x, self.norm_layer.weight, self.norm_layer.bias, self.norm_layer.eps)
        w, b = self._pair.get(self.linear_hs, self.linear_hr)
        y = ops.linear(xn.reshape(B * N, -1), w, b, precision=PRECISION)
        hs, hr = y[:, :self.c_out], y[:, self.c_out:]
        nn_ = None if num_neighbours is None else num_neighbours.reshape(B, N).float()
        return ops.graph_mean_aggregate(hs.unflatten(0, (B, N)), hr.unflatten(0, (B, N)), adj, self.num_edges, nn_,
                                        activation=activation)

    def _forward_dense(self, x, adj, num_neighbours, activation):
        """The same function as dense differentiable torch algebra - kept as an independent check of the kernels' gradients
        (tests/test_gpu_graph.py); not on any product path."""
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
        B, N = x.shape[0], x.shape[1]
        width = self.c_out_per_head * self.num_heads
        if _grad_mode(x, self.linear_hs.weight, self.attn_weight):
            # training: the logits stay extra columns of the fused projection (differentiable in attn_weight through the
            # construction of the fused weight), aggregation and its backward are one kernel each
            H, wr = self.num_heads, width * (self.num_edges + 1)
            xn = _layernorm(x, self.norm_layer)
            w, b = _FusedPair.build(self.linear_hs, self.linear_hr, self.attn_weight)
            y = _TCLinearFn.apply(xn.reshape(B * N, -1), w, b, PRECISION)
            cfg = dict(B=B, N=N, E=self.num_edges, H=H, Dh=self.c_out_per_head, mode=1, off_hr=width, off_ss=width + wr,
                       off_sr=width + wr + H, slope=self.leaky_relu.negative_slope, activation="gelu")
            att = GF.graph_aggregate(y, adj, cfg)
            return _linear(att, self.output_projection[1], activation=activation)
        xn = ops.layernorm(x, self.norm_layer.weight, self.norm_layer.bias, self.norm_layer.eps)
        w, b = self._pair.get(self.linear_hs, self.linear_hr, self.attn_weight)
        y = ops.linear(xn.reshape(B * N, -1), w, b, precision=PRECISION)
        H, wr = self.num_heads, width * (self.num_edges + 1)
        scores = (y[:, width + wr:width + wr + H], y[:, width + wr + H:width + wr + H + (self.num_edges + 1) * H])
        att = ops.graph_attention_aggregate(y[:, :width].unflatten(0, (B, N)), y[:, width:width + wr].unflatten(0, (B, N)),
                                            self.attn_weight, adj, self.num_edges, scores=scores,
                                            leaky_slope=self.leaky_relu.negative_slope, activation="gelu")
        return ops.linear(att, self.output_projection[1].weight, self.output_projection[1].bias, precision=PRECISION,
                          activation=activation)

    def _forward_dense(self, x, adj, activation):
        """Same mathematics as a dense masked softmax over all node pairs in differentiable torch algebra - an independent
        check of the kernels' gradients (tests/test_gpu_graph.py); not on any product path."""
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
        return GF.skip_gate(orig, self._skip(feat), self.config)


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


# =====================================================================================================================
# Edge-GNN (reference layers/networks/graph_layers.py:242-820): the coupling network of GraphCNF's node+edge flows
# =====================================================================================================================
class PairContext:
    """Index bookkeeping of one batch of node-pair lists, built once per ``mask_valid`` tensor and shared by all layers
    (and all couplings that receive the same tensor): ``flat_indices`` [R] = b*P + p of the valid pairs
    (graph_layers.py:339-347), ``rev`` [B,P] = 1 + compact row or 0 (:349-355), the two node rows of every compact pair.
    Building it costs the one host synchronisation of ``nonzero``; the reference pays one per EdgeGNN call (:802) plus one
    per sparse layer (:508,:650)."""

    def __init__(self, x_indices, mask_valid, num_nodes):
        B, P = mask_valid.shape
        self.B, self.P, self.N = B, P, num_nodes
        self.x_indices = (x_indices[0].contiguous(), x_indices[1].contiguous())
        flat_mask = (mask_valid.reshape(-1) == 1.0)
        self.flat_indices = torch.nonzero(flat_mask).reshape(-1)
        fm = flat_mask.long()
        self.rev = (fm * fm.cumsum(dim=0)).reshape(B, P)
        graph = torch.div(self.flat_indices, P, rounding_mode="floor")
        pair = self.flat_indices - graph * P
        self.node1 = graph * num_nodes + self.x_indices[0][pair]      # rows into [B*N, *]
        self.node2 = graph * num_nodes + self.x_indices[1][pair]
        self.R = int(self.flat_indices.numel())

    @staticmethod
    def of(x_indices, mask_valid, num_nodes):
        ctx = getattr(mask_valid, "_cnf_pair_ctx", None)
        key = (mask_valid._version, num_nodes)
        if ctx is None or ctx[0] != key:
            ctx = (key, PairContext(x_indices, mask_valid, num_nodes))
            mask_valid._cnf_pair_ctx = ctx
        return ctx[1]

    def compact(self, edge_feat):
        return edge_feat.reshape(self.B * self.P, -1).index_select(0, self.flat_indices)

    def expand(self, edge_rows):
        out = edge_rows.new_zeros(self.B * self.P, edge_rows.shape[-1])
        return out.index_copy(0, self.flat_indices, edge_rows).reshape(self.B, self.P, -1)


class StaticPairContext(PairContext):
    """A PairContext over preallocated buffers with a FIXED number of compact rows, for CUDA-graph replay: ``load`` refills the
    buffers for a new ``mask_valid`` (the one host synchronisation, outside the graph).  Rows beyond the real count are
    padding: they are gathered from / scattered to an extra all-zero slot behind the last pair (``scatter_indices``), so they
    carry zeros in, their results go nowhere and they receive zero gradients; for the kernels that look a pair's nodes up
    (``flat_indices``, ``node1/2``) they alias the first valid pair, and the edge -> node aggregation never sees them (it
    walks ``rev``)."""

    def __init__(self, x_indices, batch_size, num_nodes, rows, device):
        P = x_indices[0].numel()
        self.B, self.P, self.N, self.R = batch_size, P, num_nodes, int(rows)
        self.x_indices = (x_indices[0].contiguous(), x_indices[1].contiguous())
        self.flat_indices = torch.zeros(self.R, dtype=torch.int64, device=device)
        self.scatter_indices = torch.zeros(self.R, dtype=torch.int64, device=device)
        self.rev = torch.zeros(batch_size, P, dtype=torch.int64, device=device)
        self.node1 = torch.zeros(self.R, dtype=torch.int64, device=device)
        self.node2 = torch.zeros(self.R, dtype=torch.int64, device=device)

    @staticmethod
    def count(mask_valid):
        return int((mask_valid == 1.0).sum().item())

    def load(self, mask_valid):
        flat_mask = (mask_valid.reshape(-1) == 1.0)
        idx = torch.nonzero(flat_mask).reshape(-1)
        r = int(idx.numel())
        if r == 0 or r > self.R:
            return False
        self.flat_indices[:r] = idx
        self.scatter_indices[:r] = idx
        if r < self.R:
            self.flat_indices[r:] = idx[0]
            self.scatter_indices[r:] = self.B * self.P
        fm = flat_mask.long()
        self.rev.copy_((fm * fm.cumsum(dim=0)).reshape(self.B, self.P))
        graph = torch.div(self.flat_indices, self.P, rounding_mode="floor")
        pair = self.flat_indices - graph * self.P
        self.node1.copy_(graph * self.N + self.x_indices[0][pair])
        self.node2.copy_(graph * self.N + self.x_indices[1][pair])
        return True

    def attach(self, mask_valid):
        """Make ``PairContext.of(x_indices, mask_valid, N)`` return this context for the tensor's current version."""
        mask_valid._cnf_pair_ctx = ((mask_valid._version, self.N), self)

    def compact(self, edge_feat):
        flat = edge_feat.reshape(self.B * self.P, -1)
        return F.pad(flat, (0, 0, 0, 1)).index_select(0, self.scatter_indices)

    def expand(self, edge_rows):
        out = edge_rows.new_zeros(self.B * self.P + 1, edge_rows.shape[-1])
        return out.index_copy(0, self.scatter_indices, edge_rows)[:self.B * self.P].reshape(self.B, self.P, -1)


def _edge_to_node_dense(ctx, node_val, edge_val, edge_logit, H, mode, q=None, k=None, scale=1.0):
    """``cnf_edge_aggregate`` as differentiable torch algebra (every valid pair sends one message in each direction) - an
    independent check of the backward kernel (tests/test_gpu_graph.py); not on any product path."""
    BN, HD = node_val.shape
    Dh = HD // H
    dst = torch.cat([ctx.node1, ctx.node2])
    src = torch.cat([ctx.node2, ctx.node1])
    ev = torch.cat([edge_val, edge_val], dim=0)
    lg = torch.cat([edge_logit, edge_logit], dim=0)                              # [2R,H]
    if mode == "qkv":
        lg = lg + (q[dst].view(-1, H, Dh) * k[src].view(-1, H, Dh)).sum(-1) * scale
        mx = lg.new_full((BN, H), -3.0e38).scatter_reduce(0, dst.unsqueeze(-1).expand(-1, H), lg, reduce="amax")
        w = torch.exp(lg - mx[dst])
        denom = lg.new_zeros(BN, H).index_add(0, dst, w)
        w = w / denom[dst]
    else:
        w = torch.sigmoid(lg)
        denom = lg.new_zeros(BN, H).index_add(0, dst, w)
        w = w / denom[dst].clamp(min=1e-5)
    msg = (ev + node_val[src]).view(-1, H, Dh) * w.unsqueeze(-1)
    return node_val.new_zeros(BN, H, Dh).index_add(0, dst, msg).view(BN, HD)


class EdgeGNNLayer(nn.Module):
    """Node update from the incident edges, then edge update from the two end nodes (graph_layers.py:242-262)."""

    def __init__(self, edge2node_layer_func, node2edge_layer_func):
        super().__init__()
        self.node2edge_layer = node2edge_layer_func()
        self.edge2node_layer = edge2node_layer_func()

    def forward(self, node_feat, edge_feat, x_indices, mask_valid, **kwargs):
        """Reference calling convention: ``edge_feat`` [B,P,He] over all pairs, invalid pairs come back zeroed."""
        ctx = PairContext.of(x_indices, mask_valid, node_feat.size(1))
        node_feat, edge_rows = self.forward_compact(node_feat, ctx.compact(edge_feat), ctx)
        return node_feat, ctx.expand(edge_rows)

    def forward_compact(self, node_feat, edge_rows, ctx):
        node_feat = self.edge2node_layer.forward_compact(node_feat, edge_rows, ctx)
        edge_rows = self.node2edge_layer.forward_compact(node_feat, edge_rows, ctx)
        return node_feat, edge_rows

    @staticmethod
    def _get_sort_indices(x_indices):
        return torch.cat([x_indices[0], x_indices[1]], dim=0).sort(0, descending=False)[1]

    @staticmethod
    def _get_node_feat_by_indices(node_feat, x_indices, dim=1):
        return node_feat.index_select(index=x_indices[0], dim=dim), node_feat.index_select(index=x_indices[1], dim=dim)


class _EdgeLayerBase(nn.Module):

    def forward(self, node_feat, edge_feat, x_indices, mask_valid, **kwargs):
        ctx = PairContext.of(x_indices, mask_valid, node_feat.size(1))
        res = self.forward_compact(node_feat, ctx.compact(edge_feat), ctx)
        return ctx.expand(res) if self._returns_edges else res


class Node2EdgePlainLayer(_EdgeLayerBase):
    """edge <- skip(edge, GELU(W_e LN(edge) + W_n LN(node_a) + W_n LN(node_b)))   (graph_layers.py:297-336)."""
    _returns_edges = True

    def __init__(self, hidden_size_nodes, hidden_size_edges, skip_config=0, dp_rate=0.0, act_fn=nn.GELU):
        super().__init__()
        self.hidden_size_nodes, self.hidden_size_edges = hidden_size_nodes, hidden_size_edges
        self.skip_layer = GNNSkipConnection(hidden_size_edges, config=skip_config, dp_rate=dp_rate)
        self.dropout = nn.Dropout(dp_rate)
        self.act_fn = act_fn()
        self.node_feat_layer = nn.Sequential(nn.LayerNorm(hidden_size_nodes), TCLinear(hidden_size_nodes, hidden_size_edges))
        self.edge_feat_layer = nn.Sequential(nn.LayerNorm(hidden_size_edges), TCLinear(hidden_size_edges, hidden_size_edges))

    def forward_compact(self, node_feat, edge_rows, ctx):
        B, N = node_feat.shape[0], node_feat.shape[1]
        node_lin = _linear(_layernorm(self.dropout(node_feat), self.node_feat_layer[0]), self.node_feat_layer[1]).reshape(B * N, -1)
        edge_lin = _linear(_layernorm(self.dropout(edge_rows), self.edge_feat_layer[0]), self.edge_feat_layer[1])
        if self.training and self.dropout.p > 0:      # dropout sits between the sum and the activation (:330)
            comb = GF.gelu(self.dropout(GF.pair_combine(edge_lin, node_lin, ctx.flat_indices, ctx.x_indices, N, activation=None)))
        elif _grad_mode(node_lin, edge_lin):
            comb = GF.pair_combine(edge_lin, node_lin, ctx.flat_indices, ctx.x_indices, N, activation="gelu")
        else:
            comb = ops.pair_combine(ctx.flat_indices, ctx.x_indices, edge_lin, node_lin, N, activation="gelu")
        return self.skip_layer(orig=edge_rows, feat=comb)


class Edge2NodeQKVAttnLayer(_EdgeLayerBase):
    """Transformer-style node update: queries / keys / values from the nodes, value offset and logit bias from the edge
    between them (graph_layers.py:388-558; its dense and sparse forward passes compute the same function)."""
    _returns_edges = False

    def __init__(self, hidden_size_nodes, hidden_size_edges, num_heads=4, dp_rate=0.0, act_fn=nn.GELU, skip_config=2):
        super().__init__()
        self.hidden_size_nodes, self.hidden_size_edges, self.num_heads = hidden_size_nodes, hidden_size_edges, num_heads
        self.hidden_size_per_head = self.hidden_size_nodes // self.num_heads
        self.dot_prod_scaling = float(self.hidden_size_per_head) ** -0.5
        width = self.num_heads * self.hidden_size_per_head
        self.node_query_key_val_layer = TCLinear(hidden_size_nodes, width * 3)
        self.edge_val_layer = TCLinear(hidden_size_edges, width)
        self.edge_adj_layer = TCLinear(hidden_size_edges, self.num_heads)
        self.output_projection = TCLinear(width + self.hidden_size_nodes, self.hidden_size_nodes)
        self.__dict__["_pair"] = _FusedPair()
        self.skip_layer = GNNSkipConnection(hidden_size_nodes, config=skip_config, input_size=self.hidden_size_nodes, dp_rate=dp_rate)
        self.dropout = nn.Dropout(dp_rate)
        self.act_fn = act_fn()
        self.node_normalization = nn.LayerNorm(hidden_size_nodes)
        self.edge_normalization = nn.LayerNorm(hidden_size_edges)

    def forward_compact(self, node_feat, edge_rows, ctx):
        B, N = node_feat.shape[0], node_feat.shape[1]
        H, width = self.num_heads, self.num_heads * self.hidden_size_per_head
        node_in = _layernorm(node_feat, self.node_normalization)
        edge_in = _layernorm(edge_rows, self.edge_normalization)
        qkv = _linear(self.dropout(node_in), self.node_query_key_val_layer).reshape(B * N, 3 * width)
        q, k, v = qkv[:, :width], qkv[:, width:2 * width], qkv[:, 2 * width:]
        if _grad_mode(edge_in, self.edge_val_layer.weight, self.edge_adj_layer.weight):
            edge_val = _linear(edge_in, self.edge_val_layer)
            edge_adj = _linear(edge_in, self.edge_adj_layer)
        else:   # both edge projections read the same rows: one GEMM over the stacked weights, consumers take column slices
            w, b = self._pair.get(self.edge_val_layer, self.edge_adj_layer)
            ev = ops.linear(edge_in, w, b, precision=PRECISION)
            edge_val, edge_adj = ev[:, :width], ev[:, width:width + H]
        if _grad_mode(qkv, edge_val, edge_adj):
            cfg = dict(N=N, H=H, Dh=self.hidden_size_per_head, mode=1, off_q=0, off_k=width, off_val=2 * width,
                       scale=self.dot_prod_scaling)
            att = GF.edge_aggregate(qkv, edge_val, edge_adj, ctx.rev, cfg)
        else:
            att = ops.edge_aggregate(ctx.rev, v, edge_val, edge_adj, H, mode="qkv", node_q=q, node_k=k, scale=self.dot_prod_scaling)
        cat = torch.cat([node_in, att.reshape(B, N, width)], dim=-1)
        if self.training and self.dropout.p > 0:
            comb = GF.gelu(self.dropout(_linear(cat, self.output_projection)))
        else:
            comb = _linear(cat, self.output_projection, activation="gelu")
        return self.skip_layer(orig=node_feat, feat=comb)


class Edge2NodeAttnLayer(_EdgeLayerBase):
    """Node update with sigmoid attention driven purely by the edges (graph_layers.py:561-699)."""
    _returns_edges = False

    def __init__(self, hidden_size_nodes, hidden_size_edges, skip_config=2, num_heads=4, dp_rate=0.0, act_fn=nn.GELU):
        super().__init__()
        self.hidden_size_nodes, self.hidden_size_edges, self.num_heads = hidden_size_nodes, hidden_size_edges, num_heads
        self.hidden_size_per_head = int(self.hidden_size_nodes // self.num_heads)
        self.hidden_size_output = self.hidden_size_per_head * self.num_heads
        self.node_feat_layer = TCLinear(hidden_size_nodes, self.hidden_size_output * 2)
        self.edge_feat_layer = TCLinear(hidden_size_edges, self.hidden_size_output)
        self.edge_logits_layer = TCLinear(hidden_size_edges, self.num_heads)
        self.__dict__["_pair"] = _FusedPair()
        self.skip_layer = GNNSkipConnection(hidden_size_nodes, config=skip_config, input_size=self.hidden_size_output)
        self.dropout = nn.Dropout(dp_rate)
        self.act_fn = act_fn()
        self.node_normalization = nn.LayerNorm(hidden_size_nodes)
        self.edge_normalization = nn.LayerNorm(hidden_size_edges)

    def forward_compact(self, node_feat, edge_rows, ctx):
        B, N = node_feat.shape[0], node_feat.shape[1]
        HO = self.hidden_size_output
        node_new = _linear(_layernorm(node_feat, self.node_normalization), self.node_feat_layer).reshape(B * N, 2 * HO)
        node_self, node_ctx = node_new[:, :HO], node_new[:, HO:]
        edge_in = _layernorm(edge_rows, self.edge_normalization)
        if _grad_mode(edge_in, self.edge_feat_layer.weight, self.edge_logits_layer.weight):
            edge_new = _linear(edge_in, self.edge_feat_layer)
            edge_logits = _linear(edge_in, self.edge_logits_layer)
        else:   # one GEMM over the stacked weights of the two edge projections
            w, b = self._pair.get(self.edge_feat_layer, self.edge_logits_layer)
            ev = ops.linear(edge_in, w, b, precision=PRECISION)
            edge_new, edge_logits = ev[:, :HO], ev[:, HO:HO + self.num_heads]
        if _grad_mode(node_new, edge_new, edge_logits):
            cfg = dict(N=N, H=self.num_heads, Dh=self.hidden_size_per_head, mode=0, off_val=HO)
            att = GF.edge_aggregate(node_new, edge_new, edge_logits, ctx.rev, cfg)
        else:
            att = ops.edge_aggregate(ctx.rev, node_ctx, edge_new, edge_logits, self.num_heads, mode="sigmoid")
        comb = GF.gelu(self.dropout(node_self + att)).reshape(B, N, HO)
        return self.skip_layer(orig=node_feat, feat=comb)


class EdgeGNN(nn.Module):
    """Input MLPs, ``num_layers`` Edge-GNN layers, output MLPs for nodes and edges (graph_layers.py:737-820).  Edge features
    live as compact rows of the valid pairs from the input MLP to the output MLP; only the final edge output is scattered
    back to [B,P,c_out_edges] (zeros at invalid pairs, as upstream :812-813)."""

    def __init__(self, c_in_nodes, c_in_edges, c_out_nodes, c_out_edges, edge_gnn_layer_func, num_layers=4, max_neighbours=-1):
        super().__init__()
        self.c_in_nodes, self.c_in_edges, self.c_out_nodes, self.c_out_edges = c_in_nodes, c_in_edges, c_out_nodes, c_out_edges
        self.layers = nn.ModuleList([edge_gnn_layer_func() for _ in range(num_layers)])
        hidden_size_edges = self.layers[0].node2edge_layer.hidden_size_edges
        hidden_size_nodes = self.layers[0].node2edge_layer.hidden_size_nodes
        self.input_layer_edges = self._create_input_network(c_in_edges, hidden_size_edges)
        self.input_layer_nodes = self._create_input_network(c_in_nodes, hidden_size_nodes)
        self.out_layer_edges = self._create_output_network(hidden_size_edges, c_out_edges)
        self.out_layer_nodes = self._create_output_network(hidden_size_nodes, c_out_nodes)
        if max_neighbours > 0:
            self.max_neighbours = max_neighbours
            self.node_neighbour_embed = TCLinear(max_neighbours + 1, hidden_size_nodes)

    def _create_input_network(self, c_in, hidden_size):
        return nn.Sequential(TCLinear(c_in, hidden_size), nn.GELU(), TCLinear(hidden_size, hidden_size))

    def _create_output_network(self, hidden_size, c_out):
        return nn.Sequential(nn.LayerNorm(hidden_size), TCLinear(hidden_size, hidden_size), nn.GELU(), TCLinear(hidden_size, c_out))

    @staticmethod
    def _mlp_in(net, x):
        return _linear(_linear(x, net[0], activation="gelu"), net[2])

    @staticmethod
    def _mlp_out(net, x):
        return _linear(_linear(_layernorm(x, net[0]), net[1], activation="gelu"), net[3])

    def forward(self, z_nodes, z_edges, length, x_indices, mask_valid, channel_padding_mask=None, binary_adjacency=None, **kwargs):
        ctx = PairContext.of(x_indices, mask_valid, z_nodes.size(1))
        nodes_feat = self._mlp_in(self.input_layer_nodes, z_nodes)
        edge_rows = self._mlp_in(self.input_layer_edges, ctx.compact(z_edges))
        if binary_adjacency is not None and hasattr(self, "node_neighbour_embed"):
            num_neighbours = binary_adjacency.sum(dim=-1).long().clamp(max=self.max_neighbours)
            nodes_feat = nodes_feat + F.embedding(num_neighbours, self.node_neighbour_embed.weight.t()) + self.node_neighbour_embed.bias
        for layer in self.layers:
            nodes_feat, edge_rows = layer.forward_compact(nodes_feat, edge_rows, ctx)
        nodes_out = self._mlp_out(self.out_layer_nodes, nodes_feat)
        edges_out = ctx.expand(self._mlp_out(self.out_layer_edges, edge_rows))
        if channel_padding_mask is not None:
            nodes_out = nodes_out * channel_padding_mask
        return nodes_out, edges_out
