"""Autograd-aware glue of the graph coupling networks (``RGCNNet``, ``EdgeGNN``): every function is one forward kernel of
``csrc/graph_ops.cu`` and, under ``torch.enable_grad()``, one backward kernel of ``csrc/graph_ops_bwd.cu``.

The aggregation functions take the *whole* projection output (``hs | hr | logits`` or ``q | k | v`` as column blocks of
one row-major matrix) plus column offsets, and return one gradient matrix of the same shape - so autograd never
materialises a zero-padded copy per column slice, and the projection's backward GEMM reads that matrix as is.
"""
from __future__ import annotations

import torch

from . import _lib as L
from . import ops
from .ops import ACTIVATION, _call, _f32, _ptr


def _needs_grad(*tensors) -> bool:
    return torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors)


def _dense(t):
    t = t if t.dtype == torch.float32 else t.float()
    return t if t.is_contiguous() else t.contiguous()


# ----------------------------------------------------------------------------------------------
# GELU
# ----------------------------------------------------------------------------------------------
def _gelu_call(x, grad_y=None):
    x = _f32(x, "x")
    y = torch.empty_like(x)
    a = L.GeluArgs()
    gy = None if grad_y is None else _f32(grad_y, "grad_y", x.shape)
    a.n, a.x, a.grad_y, a.y = x.numel(), _ptr(x), _ptr(gy), _ptr(y)
    _call("cnf_gelu", a, x, (x, gy))
    return y


class _Gelu(torch.autograd.Function):

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return _gelu_call(x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return _gelu_call(x, g)


def gelu(x):
    """erf-GELU (``nn.GELU``) where it is not the epilogue of a projection."""
    return _Gelu.apply(x) if _needs_grad(x) else _gelu_call(x)


# ----------------------------------------------------------------------------------------------
# LayerNorm
# ----------------------------------------------------------------------------------------------
class _LayerNorm(torch.autograd.Function):

    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        ctx.eps = eps
        x = _dense(x)
        ctx.save_for_backward(x, weight)
        return ops.layernorm(x, weight, bias, eps)

    @staticmethod
    def backward(ctx, g):
        x, weight = ctx.saved_tensors
        g = _f32(g, "grad_y", x.shape)
        H = x.shape[-1]
        gx = torch.empty_like(x)
        need_p = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        gg = torch.zeros(H, dtype=torch.float32, device=x.device) if need_p else None
        gb = torch.zeros(H, dtype=torch.float32, device=x.device) if need_p else None
        w = _f32(weight, "weight", (H,))
        a = L.LayernormBwdArgs()
        a.M, a.H, a.x, a.gamma, a.eps, a.grad_y = x.numel() // H, H, _ptr(x), _ptr(w), float(ctx.eps), _ptr(g)
        a.grad_x, a.grad_gamma, a.grad_beta = _ptr(gx), _ptr(gg), _ptr(gb)
        _call("cnf_layernorm_bwd", a, x, (x, w, g))
        return gx, gg, gb, None


def layernorm(x, weight, bias, eps=1e-5):
    if _needs_grad(x, weight, bias):
        return _LayerNorm.apply(x, weight, bias, float(eps))
    return ops.layernorm(x, weight, bias, eps)


# ----------------------------------------------------------------------------------------------
# GNNSkipConnection
# ----------------------------------------------------------------------------------------------
class _SkipGate(torch.autograd.Function):

    @staticmethod
    def forward(ctx, orig, skip, config):
        orig, skip = _dense(orig), _dense(skip)
        ctx.config = config
        ctx.save_for_backward(orig, skip)
        return ops.skip_gate(orig, skip, config)

    @staticmethod
    def backward(ctx, g):
        orig, skip = ctx.saved_tensors
        g = _f32(g, "grad_out", orig.shape)
        H = orig.shape[-1]
        go, gs = torch.empty_like(orig), torch.empty_like(skip)
        a = L.SkipGateBwdArgs()
        a.M, a.H, a.config = orig.numel() // H, H, int(ctx.config)
        a.orig, a.skip, a.grad_out, a.grad_orig, a.grad_skip = _ptr(orig), _ptr(skip), _ptr(g), _ptr(go), _ptr(gs)
        _call("cnf_skip_gate_bwd", a, orig, (orig, skip, g))
        return go, gs, None


def skip_gate(orig, skip, config):
    if _needs_grad(orig, skip):
        return _SkipGate.apply(orig, skip, int(config))
    return ops.skip_gate(orig, skip, config)


# ----------------------------------------------------------------------------------------------
# RelationGraphConv / RelationGraphAttention neighbour aggregation on the fused projection output
# ----------------------------------------------------------------------------------------------
def _agg_args(y, adjacency, cfg, num_neighbours):
    """cnf_graph_aggregate_args for the projection output ``y`` [B*N, ld] with column blocks described by ``cfg``."""
    B, N, E, H, Dh = cfg["B"], cfg["N"], cfg["E"], cfg["H"], cfg["Dh"]
    ld = y.stride(0)
    base = y.data_ptr()
    g = L.GraphAggregateArgs()
    g.B, g.N, g.E, g.H, g.Dh = B, N, E, H, Dh
    g.adjacency, g.ld_hs, g.ld_hr = _ptr(adjacency), ld, ld
    g.hr = base + 4 * cfg["off_hr"]
    if cfg["mode"] == 1:
        g.hs = None
        g.score_s, g.score_r = base + 4 * cfg["off_ss"], base + 4 * cfg["off_sr"]
        g.ld_score_s, g.ld_score_r = ld, ld
        g.num_neighbours = None
    else:
        g.hs = base + 4 * cfg["off_hs"]
        g.score_s, g.score_r, g.ld_score_s, g.ld_score_r = None, None, 0, 0
        g.num_neighbours = _ptr(num_neighbours)
    g.mode, g.leaky_slope, g.activation = cfg["mode"], float(cfg.get("slope", 0.0)), ACTIVATION[cfg.get("activation")]
    return g


class _GraphAggregate(torch.autograd.Function):

    @staticmethod
    def forward(ctx, y, adjacency, num_neighbours, cfg):
        y = _dense(y)
        g = _agg_args(y, adjacency, cfg, num_neighbours)
        out = torch.empty(cfg["B"], cfg["N"], cfg["H"] * cfg["Dh"], dtype=torch.float32, device=y.device)
        g.out = _ptr(out)
        _call("cnf_graph_aggregate", g, y, (y, adjacency, num_neighbours))
        ctx.cfg = cfg
        ctx.save_for_backward(y, adjacency, num_neighbours)
        return out

    @staticmethod
    def backward(ctx, g_out):
        y, adjacency, num_neighbours = ctx.saved_tensors
        cfg = ctx.cfg
        g_out = _f32(g_out, "grad_out")
        gy = torch.zeros_like(y)
        ld, base = gy.stride(0), gy.data_ptr()
        a = L.GraphAggregateBwdArgs()
        a.fwd = _agg_args(y, adjacency, cfg, num_neighbours)
        a.grad_out = _ptr(g_out)
        a.grad_hr, a.ld_grad_hr = base + 4 * cfg["off_hr"], ld
        if cfg["mode"] == 1:
            a.grad_score_s, a.grad_score_r = base + 4 * cfg["off_ss"], base + 4 * cfg["off_sr"]
            a.ld_grad_score_s, a.ld_grad_score_r = ld, ld
        else:
            a.grad_hs, a.ld_grad_hs = base + 4 * cfg["off_hs"], ld
        _call("cnf_graph_aggregate_bwd", a, y, (y, adjacency, num_neighbours, g_out, gy))
        return gy, None, None, None


def graph_aggregate(y, adjacency, cfg, num_neighbours=None):
    """``y`` [B*N, ld]: the projection output with column blocks ``cfg["off_*"]``; ``adjacency`` [B,N,N] int64.
    mode 1 (attention): blocks hr [(E+1)*H*Dh], score_s [H], score_r [(E+1)*H]; mode 0 (mean): hs [Dh], hr [E*Dh]."""
    adjacency = ops._adjacency(adjacency, cfg["B"], cfg["N"])
    if num_neighbours is not None:
        num_neighbours = _f32(num_neighbours, "num_neighbours").reshape(cfg["B"], cfg["N"])
    return _GraphAggregate.apply(y, adjacency, num_neighbours, cfg)


# ----------------------------------------------------------------------------------------------
# Edge-GNN: edges -> nodes attention, nodes -> edges combine
# ----------------------------------------------------------------------------------------------
def _edge_args(node_mat, edge_val, edge_logit, rev, cfg):
    B, P = rev.shape
    a = L.EdgeAggregateArgs()
    a.B, a.N, a.H, a.Dh, a.R = B, cfg["N"], cfg["H"], cfg["Dh"], edge_val.shape[0]
    base, ld = node_mat.data_ptr(), node_mat.stride(0)
    a.rev = _ptr(rev)
    a.node_val, a.ld_node_val = base + 4 * cfg["off_val"], ld
    if cfg["mode"] == 1:
        a.node_q, a.node_k, a.ld_node_q, a.ld_node_k = base + 4 * cfg["off_q"], base + 4 * cfg["off_k"], ld, ld
    a.edge_val, a.ld_edge_val = _ptr(edge_val), edge_val.stride(0)
    a.edge_logit, a.ld_edge_logit = _ptr(edge_logit), edge_logit.stride(0)
    a.mode, a.scale = cfg["mode"], float(cfg.get("scale", 1.0))
    return a


class _EdgeAggregate(torch.autograd.Function):

    @staticmethod
    def forward(ctx, node_mat, edge_val, edge_logit, rev, cfg):
        node_mat, edge_val, edge_logit = _dense(node_mat), _dense(edge_val), _dense(edge_logit)
        a = _edge_args(node_mat, edge_val, edge_logit, rev, cfg)
        out = torch.empty(node_mat.shape[0], cfg["H"] * cfg["Dh"], dtype=torch.float32, device=node_mat.device)
        a.out = _ptr(out)
        _call("cnf_edge_aggregate", a, node_mat, (node_mat, edge_val, edge_logit, rev))
        ctx.cfg = cfg
        ctx.save_for_backward(node_mat, edge_val, edge_logit, rev)
        return out

    @staticmethod
    def backward(ctx, g_out):
        node_mat, edge_val, edge_logit, rev = ctx.saved_tensors
        cfg = ctx.cfg
        g_out = _f32(g_out, "grad_out")
        gn, gev, gel = torch.zeros_like(node_mat), torch.zeros_like(edge_val), torch.zeros_like(edge_logit)
        a = L.EdgeAggregateBwdArgs()
        a.fwd = _edge_args(node_mat, edge_val, edge_logit, rev, cfg)
        base, ld = gn.data_ptr(), gn.stride(0)
        a.grad_out = _ptr(g_out)
        a.grad_node_val, a.ld_grad_node_val = base + 4 * cfg["off_val"], ld
        if cfg["mode"] == 1:
            a.grad_node_q, a.grad_node_k = base + 4 * cfg["off_q"], base + 4 * cfg["off_k"]
            a.ld_grad_node_q, a.ld_grad_node_k = ld, ld
        a.grad_edge_val, a.ld_grad_edge_val = _ptr(gev), gev.stride(0)
        a.grad_edge_logit, a.ld_grad_edge_logit = _ptr(gel), gel.stride(0)
        _call("cnf_edge_aggregate_bwd", a, node_mat, (node_mat, edge_val, edge_logit, rev, g_out, gn, gev, gel))
        return gn, gev, gel, None, None


def edge_aggregate(node_mat, edge_val, edge_logit, rev, cfg):
    """``node_mat`` [B*N, ld] with the value (and query / key) column blocks at ``cfg["off_val"|"off_q"|"off_k"]``;
    ``edge_val`` [R, H*Dh], ``edge_logit`` [R, H] compact pair rows; ``rev`` [B,P].  Returns [B*N, H*Dh]."""
    return _EdgeAggregate.apply(node_mat, edge_val, edge_logit, rev.contiguous(), cfg)


def _pair_args(edge_lin, node_lin, flat_indices, x1, x2, num_nodes, activation):
    a = L.PairCombineArgs()
    a.R, a.N, a.He = edge_lin.shape[0], int(num_nodes), edge_lin.shape[-1]
    a.flat_indices, a.x_indices1, a.x_indices2 = _ptr(flat_indices), _ptr(x1), _ptr(x2)
    a.edge_lin, a.node_lin, a.ld_edge, a.ld_node = _ptr(edge_lin), _ptr(node_lin), edge_lin.stride(0), node_lin.stride(0)
    a.activation = ACTIVATION[activation]
    return a


class _PairCombine(torch.autograd.Function):

    @staticmethod
    def forward(ctx, edge_lin, node_lin, flat_indices, x1, x2, num_nodes, activation):
        edge_lin, node_lin = _dense(edge_lin), _dense(node_lin)
        out = torch.empty_like(edge_lin)
        if edge_lin.shape[0] > 0:
            a = _pair_args(edge_lin, node_lin, flat_indices, x1, x2, num_nodes, activation)
            a.out = _ptr(out)
            _call("cnf_pair_combine", a, node_lin, (edge_lin, node_lin, flat_indices, x1, x2))
        ctx.num_nodes, ctx.activation = num_nodes, activation
        ctx.save_for_backward(edge_lin, node_lin, flat_indices, x1, x2)
        return out

    @staticmethod
    def backward(ctx, g_out):
        edge_lin, node_lin, flat_indices, x1, x2 = ctx.saved_tensors
        g_out = _f32(g_out, "grad_out", edge_lin.shape)
        ge, gn = torch.empty_like(edge_lin), torch.zeros_like(node_lin)
        if edge_lin.shape[0] > 0:
            a = L.PairCombineBwdArgs()
            a.fwd = _pair_args(edge_lin, node_lin, flat_indices, x1, x2, ctx.num_nodes, ctx.activation)
            a.grad_out, a.grad_edge_lin, a.grad_node_lin = _ptr(g_out), _ptr(ge), _ptr(gn)
            a.ld_grad_edge, a.ld_grad_node = ge.stride(0), gn.stride(0)
            _call("cnf_pair_combine_bwd", a, node_lin, (edge_lin, node_lin, flat_indices, x1, x2, g_out, ge, gn))
        return ge, gn, None, None, None, None, None


def pair_combine(edge_lin, node_lin, flat_indices, x_indices, num_nodes, activation="gelu"):
    """``act(edge_lin[r] + node_lin[b, x1[p]] + node_lin[b, x2[p]])`` for every compact pair row (flat_indices[r] = b*P + p)."""
    return _PairCombine.apply(edge_lin, node_lin.reshape(-1, node_lin.shape[-1]), flat_indices.contiguous(),
                              x_indices[0].contiguous(), x_indices[1].contiguous(), int(num_nodes), activation)
