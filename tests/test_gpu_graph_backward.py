"""Backward kernels of the graph coupling networks' glue (csrc/graph_ops_bwd.cu) against float64 autograd of the same
functions written as plain torch algebra, and the Edge-GNN training step (reference general/train.py:148-152 differentiates
EdgeGNN through autograd) against autograd through the CPU oracle."""
import pytest
import torch
import torch.nn.functional as F

from conftest import assert_close, load_golden
from oracle import graph_oracle as GO

pytestmark = pytest.mark.gpu


def _graphs(gen, B, N, E, p=0.25, symmetric=True):
    length = torch.randint(max(2, N // 2), N + 1, (B,), generator=gen)
    up = (torch.rand(B, N, N, generator=gen) < p).long() * torch.randint(1, E + 1, (B, N, N), generator=gen)
    if symmetric:
        up = torch.triu(up, diagonal=1)
        up = up + up.transpose(1, 2)
    else:
        up = up * (1 - torch.eye(N, dtype=torch.long))
    valid = torch.arange(N)[None, :] < length[:, None]
    return up * (valid[:, :, None] & valid[:, None, :]).long(), length


def test_gelu_forward_backward():
    from categoricalnf_b200 import graph_functional as GF
    x = torch.randn(1000, 37, generator=torch.Generator().manual_seed(0)) * 3
    g = torch.randn(1000, 37, generator=torch.Generator().manual_seed(1))
    xr = x.double().requires_grad_(True)
    yr = F.gelu(xr)
    (yr * g).sum().backward()
    xc = x.cuda().requires_grad_(True)
    y = GF.gelu(xc)
    (y * g.cuda()).sum().backward()
    assert_close(y, yr, rtol=1e-5, atol=1e-6, what="gelu")
    assert_close(xc.grad, xr.grad, rtol=1e-5, atol=1e-6, what="gelu'")


@pytest.mark.parametrize("M,H", [(37, 32), (1000, 384), (5, 30), (4096, 192)])
def test_layernorm_backward(M, H):
    from categoricalnf_b200 import graph_functional as GF
    gen = torch.Generator().manual_seed(M)
    x, w, b = torch.randn(M, H, generator=gen) * 2 + 0.5, torch.randn(H, generator=gen), torch.randn(H, generator=gen)
    g = torch.randn(M, H, generator=gen)
    xr, wr, br = (t.double().requires_grad_(True) for t in (x, w, b))
    (F.layer_norm(xr, (H,), wr, br, 1e-5) * g).sum().backward()
    xc, wc, bc = (t.cuda().requires_grad_(True) for t in (x, w, b))
    (GF.layernorm(xc, wc, bc, 1e-5) * g.cuda()).sum().backward()
    assert_close(xc.grad, xr.grad, rtol=1e-4, atol=2e-5, what="grad x")
    assert_close(wc.grad, wr.grad, rtol=1e-4, atol=1e-4 * float(wr.grad.abs().max()), what="grad gamma")
    assert_close(bc.grad, br.grad, rtol=1e-4, atol=1e-4 * float(br.grad.abs().max()), what="grad beta")


@pytest.mark.parametrize("config", [0, 1, 2])
def test_skip_gate_backward(config):
    from categoricalnf_b200 import graph_functional as GF
    gen = torch.Generator().manual_seed(config)
    M, H = 333, 48
    orig, skip = torch.randn(M, H, generator=gen), torch.randn(M, H * (1 if config == 0 else 2), generator=gen)
    g = torch.randn(M, H, generator=gen)
    o, s = orig.double().requires_grad_(True), skip.double().requires_grad_(True)
    if config == 0:
        ref = o + s
    else:
        val, gate = s[:, :H], torch.sigmoid(s[:, H:])
        ref = o + val * gate if config == 1 else o * (1 - gate) + val * gate
    (ref * g).sum().backward()
    oc, sc = orig.cuda().requires_grad_(True), skip.cuda().requires_grad_(True)
    (GF.skip_gate(oc, sc, config) * g.cuda()).sum().backward()
    assert_close(oc.grad, o.grad, rtol=1e-5, atol=1e-6, what="grad orig")
    assert_close(sc.grad, s.grad, rtol=1e-5, atol=1e-6, what="grad skip")


@pytest.mark.parametrize("B,N,E,C,act,given_nn", [(3, 9, 3, 8, None, True), (5, 20, 1, 32, "gelu", False), (2, 38, 3, 30, "gelu", True)])
def test_mean_aggregate_backward(B, N, E, C, act, given_nn):
    """RelationGraphConv aggregation (graph_layers.py:37-50) on a non-symmetric adjacency: the sum runs over the first index."""
    from categoricalnf_b200 import graph_functional as GF
    gen = torch.Generator().manual_seed(B * N)
    adj, _ = _graphs(gen, B, N, E, symmetric=False)
    y = torch.randn(B * N, C * (1 + E), generator=gen)
    g = torch.randn(B, N, C, generator=gen)
    nn_ = (adj > 0).sum(dim=1).float().clamp(max=2) if given_nn else None
    yr = y.double().requires_grad_(True)
    hs, hr = yr[:, :C].view(B, N, C), yr[:, C:].view(B, N, E, C)
    onehot = F.one_hot(adj, E + 1)[..., 1:].double()                                   # [B,j,i,E]
    n = onehot.sum(dim=[1, 3]) if nn_ is None else nn_.double()
    ref = hs + torch.einsum("bjie,bjec->bic", onehot, hr) / n.unsqueeze(-1).clamp(min=1e-5)
    ref = F.gelu(ref) if act else ref
    (ref * g).sum().backward()
    yc = y.cuda().requires_grad_(True)
    cfg = dict(B=B, N=N, E=E, H=1, Dh=C, mode=0, off_hs=0, off_hr=C, activation=act)
    out = GF.graph_aggregate(yc, adj.cuda(), cfg, None if nn_ is None else nn_.cuda())
    (out * g.cuda()).sum().backward()
    assert_close(out, ref, rtol=1e-5, atol=1e-5, what="forward")
    assert_close(yc.grad, yr.grad, rtol=1e-4, atol=1e-5, what="grad y")


@pytest.mark.parametrize("B,N,E,H,Dh,act", [(3, 9, 1, 4, 8, "gelu"), (4, 20, 3, 4, 16, "gelu"), (2, 38, 2, 2, 6, None)])
def test_attention_aggregate_backward(B, N, E, H, Dh, act):
    """RelationGraphAttention aggregation (graph_layers.py:92-154) with the logits as columns of the projection output."""
    from categoricalnf_b200 import graph_functional as GF
    gen = torch.Generator().manual_seed(B + N)
    adj, _ = _graphs(gen, B, N, E, symmetric=False)
    width = H * Dh
    wr = width * (E + 1)
    cols = width + wr + H + (E + 1) * H
    cols += (-cols) % 4
    y = torch.randn(B * N, cols, generator=gen)
    g = torch.randn(B, N, width, generator=gen)
    yr = y.double().requires_grad_(True)
    hr = yr[:, width:width + wr].view(B, N, E + 1, H, Dh)
    ss = yr[:, width + wr:width + wr + H].view(B, N, H)
    sr = yr[:, width + wr + H:width + wr + H + (E + 1) * H].view(B, N, E + 1, H)
    eye = torch.eye(N, dtype=torch.long).unsqueeze(0)
    etype = torch.where(eye.bool(), torch.full_like(adj, E + 1), adj)                 # [B,i,j]
    slot = (etype - 1).clamp(min=0)
    sr_ij = torch.gather(sr.unsqueeze(1).expand(B, N, N, E + 1, H), 3, slot.view(B, N, N, 1, 1).expand(B, N, N, 1, H)).squeeze(3)
    logits = F.leaky_relu(ss.unsqueeze(2) + sr_ij, 0.2).masked_fill(~(etype > 0).unsqueeze(-1), -9e15)
    probs = torch.softmax(logits, dim=2)
    hr_ij = torch.gather(hr.unsqueeze(1).expand(B, N, N, E + 1, H, Dh), 3,
                         slot.view(B, N, N, 1, 1, 1).expand(B, N, N, 1, H, Dh)).squeeze(3)
    ref = (probs.unsqueeze(-1) * hr_ij).sum(dim=2).reshape(B, N, width)
    ref = F.gelu(ref) if act else ref
    (ref * g).sum().backward()
    yc = y.cuda().requires_grad_(True)
    cfg = dict(B=B, N=N, E=E, H=H, Dh=Dh, mode=1, off_hr=width, off_ss=width + wr, off_sr=width + wr + H, slope=0.2, activation=act)
    out = GF.graph_aggregate(yc, adj.cuda(), cfg)
    (out * g.cuda()).sum().backward()
    assert_close(out, ref, rtol=1e-4, atol=1e-5, what="forward")
    assert_close(yc.grad, yr.grad, rtol=1e-4, atol=2e-5, what="grad y")


def _pairs(N):
    i1 = torch.tensor([i for i in range(N) for j in range(i + 1, N)])
    i2 = torch.tensor([j for i in range(N) for j in range(i + 1, N)])
    return i1, i2


@pytest.mark.parametrize("mode,B,N,H,Dh", [("sigmoid", 3, 7, 4, 8), ("qkv", 3, 9, 4, 8), ("sigmoid", 2, 38, 2, 6), ("qkv", 4, 20, 4, 16)])
def test_edge_aggregate_and_pair_combine_backward(mode, B, N, H, Dh):
    from categoricalnf_b200 import graph_functional as GF
    from categoricalnf_b200.layers.networks.graph_layers import PairContext, _edge_to_node_dense
    gen = torch.Generator().manual_seed(N + H)
    i1, i2 = _pairs(N)
    length = torch.randint(2, N + 1, (B,), generator=gen)
    mask_valid = ((i1[None, :] < length[:, None]) & (i2[None, :] < length[:, None])).float()
    mask_valid = mask_valid * (torch.rand(B, i1.numel(), generator=gen) < 0.6).float()
    mask_valid[0, :] = 0                                                              # a graph without valid pairs
    ctx = PairContext((i1.cuda(), i2.cuda()), mask_valid.cuda(), N)
    HD, R = H * Dh, ctx.R
    ncols = 3 * HD if mode == "qkv" else 2 * HD
    node_mat = torch.randn(B * N, ncols, generator=gen)
    ev, el = torch.randn(R, HD, generator=gen), torch.randn(R, H, generator=gen) * 2
    if mode == "sigmoid":
        el[: max(1, R // 4)] -= 30.0                                                  # some weight sums fall under the 1e-5 clamp
    g = torch.randn(B * N, HD, generator=gen)
    # float64 reference on the GPU (index bookkeeping of PairContext lives there)
    nm, e1, e2 = (t.double().cuda().requires_grad_(True) for t in (node_mat, ev, el))
    if mode == "qkv":
        ref = _edge_to_node_dense(ctx, nm[:, 2 * HD:], e1, e2, H, "qkv", nm[:, :HD], nm[:, HD:2 * HD], Dh ** -0.5)
        cfg = dict(N=N, H=H, Dh=Dh, mode=1, off_q=0, off_k=HD, off_val=2 * HD, scale=Dh ** -0.5)
    else:
        ref = _edge_to_node_dense(ctx, nm[:, HD:], e1, e2, H, "sigmoid")
        cfg = dict(N=N, H=H, Dh=Dh, mode=0, off_val=HD)
    (ref * g.double().cuda()).sum().backward()
    nc, ec, lc = (t.cuda().requires_grad_(True) for t in (node_mat, ev, el))
    out = GF.edge_aggregate(nc, ec, lc, ctx.rev, cfg)
    (out * g.cuda()).sum().backward()
    assert_close(out, ref, rtol=1e-4, atol=1e-5, what="forward")
    assert_close(nc.grad, nm.grad, rtol=1e-4, atol=2e-5, what="grad node matrix")
    assert_close(ec.grad, e1.grad, rtol=1e-4, atol=2e-5, what="grad edge values")
    assert_close(lc.grad, e2.grad, rtol=1e-4, atol=2e-5, what="grad edge logits")

    He = 12
    edge_lin, node_lin = torch.randn(R, He, generator=gen), torch.randn(B * N, He, generator=gen)
    gp = torch.randn(R, He, generator=gen)
    for act in ("gelu", None):
        er, nr = edge_lin.double().cuda().requires_grad_(True), node_lin.double().cuda().requires_grad_(True)
        ref = er + nr[ctx.node1] + nr[ctx.node2]
        ref = F.gelu(ref) if act else ref
        (ref * gp.double().cuda()).sum().backward()
        ec2, nc2 = edge_lin.cuda().requires_grad_(True), node_lin.cuda().requires_grad_(True)
        out = GF.pair_combine(ec2, nc2, ctx.flat_indices, ctx.x_indices, N, activation=act)
        (out * gp.cuda()).sum().backward()
        assert_close(out, ref, rtol=1e-5, atol=1e-5, what="pair_combine forward")
        assert_close(ec2.grad, er.grad, rtol=1e-4, atol=1e-5, what="grad edge_lin")
        assert_close(nc2.grad, nr.grad, rtol=1e-4, atol=1e-5, what="grad node_lin")


@pytest.mark.parametrize("name", ["edge_gnn_attn_dense", "edge_gnn_qkv_dense"])
def test_edge_gnn_training_gradients_vs_oracle(name):
    """Edge-GNN under autograd: output equal to the golden, parameter and input gradients equal to autograd through the
    CPU oracle - every glue op differentiated by a kernel of graph_ops_bwd.cu."""
    from test_gpu_graph import _build_edge_gnn
    from categoricalnf_b200 import ops
    g = load_golden(name)
    net = _build_edge_gnn(g).train()
    sd = {k[len("sd__"):]: v.clone().requires_grad_(v.is_floating_point()) for k, v in g.items() if k.startswith("sd__")}
    zn, ze = g.z_nodes.clone().requires_grad_(True), g.z_edges.clone().requires_grad_(True)
    ref_n, ref_e = GO.edge_gnn(sd, zn, ze, (g.x_indices1, g.x_indices2), g.mask_valid, num_layers=g.layers, qkv=bool(g.qkv), pad=g.pad,
                               binary_adjacency=None, max_neighbours=g.max_neighbours)
    gen = torch.Generator().manual_seed(4)
    wn, we = torch.randn(ref_n.shape, generator=gen), torch.randn(ref_e.shape, generator=gen)
    ((ref_n * wn).sum() + (ref_e * we).sum()).backward()
    znc, zec = g.z_nodes.cuda().requires_grad_(True), g.z_edges.cuda().requires_grad_(True)
    before = ops.launch_count()
    nodes, edges = net(znc, zec, length=g.length.cuda(), x_indices=(g.x_indices1.cuda(), g.x_indices2.cuda()),
                       mask_valid=g.mask_valid.cuda(), channel_padding_mask=g.pad.cuda(), binary_adjacency=None)
    ((nodes * wn.cuda()).sum() + (edges * we.cuda()).sum()).backward()
    assert ops.launch_count() - before > 40, "the training step must run on the C-ABI kernels"
    assert_close(nodes, g.nodes_out, rtol=1e-4, atol=2e-5, what="nodes_out")
    assert_close(znc.grad, zn.grad, rtol=1e-3, atol=2e-5, what="grad z_nodes")
    assert_close(zec.grad, ze.grad, rtol=1e-3, atol=2e-5, what="grad z_edges")
    # (the bias of the logit layer of the query-key attention has a zero gradient - softmax is shift invariant - so a
    #  per-parameter scale would compare rounding noise; floor it by the largest gradient of the network)
    floor = 1e-3 * max(float(v.grad.abs().max()) for v in sd.values() if v.grad is not None)
    for k, p in net.named_parameters():
        if sd[k].grad is None:
            continue
        scale = max(float(sd[k].grad.abs().max()), floor)
        assert p.grad is not None, k
        assert_close(p.grad / scale, sd[k].grad / scale, rtol=1e-3, atol=3e-4, what="grad " + k)


def test_graphcnf_training_step_gradients_vs_reference():
    """One training step of GraphCNF (BASELINE config 4 in small): the forward of the golden run with autograd enabled, the
    loss the golden script used, and the gradient of EVERY parameter against the gradients the unmodified reference produced
    for it (tests/golden/graphcnf_small_grads.npz) - encodings, node / edge-attribute / virtual-edge flows, RGCN and Edge-GNN
    coupling networks all differentiate through the backward kernels."""
    from test_gpu_graph import _build_graphcnf, _sd
    from categoricalnf_b200 import ops
    g, gg = load_golden("graphcnf_small"), load_golden("graphcnf_small_grads")
    model = _build_graphcnf(g.N, sd=_sd(g)).train()
    before = ops.launch_count()
    z, ldj = model(g.x.cuda(), adjacency=g.adjacency.cuda(), length=g.length.cuda(), u_noise=g.u_nodes.cuda(),
                   u_noise_edges=g.u_edges.cuda(), u_noise_virtual=g.u_virtual.cuda())
    assert_close(z, gg.z_train, what="z nodes")
    assert_close(ldj, gg.ldj_train, rtol=1e-4, atol=5e-4, what="ldj")
    (-(ldj.sum()) + (z * gg.wz.cuda()).sum()).backward()
    assert ops.launch_count() - before > 300, "forward + backward must run on the C-ABI kernels"
    ref = {k[len("grad__"):]: v for k, v in gg.items() if k.startswith("grad__")}
    floor = 1e-3 * max(float(v.abs().max()) for v in ref.values())
    checked = 0
    for name, p in model.named_parameters():
        if name not in ref:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, "unexpected gradient for " + name
            continue
        assert p.grad is not None, "no gradient for " + name
        scale = max(float(ref[name].abs().max()), floor)
        assert_close(p.grad / scale, ref[name] / scale, rtol=2e-3, atol=5e-4, what="grad " + name)
        checked += 1
    assert checked == len(ref)


def test_graphcnf_training_step_cuda_graph_replay():
    """GraphedTrainingStep: forward + backward replayed from a CUDA graph leave the same loss and parameter gradients as the
    eager training step on the same noise - on the capture batch and on later batches that only replay."""
    from test_gpu_graph import _build_graphcnf, _graphs, _sd
    from categoricalnf_b200.experiments.molecule_generation import GraphedTrainingStep
    g = load_golden("graphcnf_small")
    model = _build_graphcnf(g.N, sd=_sd(g)).train()
    loss_fn = lambda z, ldj, length: -(ldj / length.to(ldj.dtype)).mean() + 0.01 * (z ** 2).mean()
    step = GraphedTrainingStep(model, loss_fn=loss_fn, bucket=16)
    gen = torch.Generator().manual_seed(9)
    B, N = g.x.shape
    P = N * (N - 1) // 2
    params = [p for p in model.parameters() if p.requires_grad]
    for trial in range(3):
        adj, length = _graphs(gen, B, N, 3, p=0.3)
        x = torch.randint(0, 5, (B, N), generator=gen) * (torch.arange(N)[None, :] < length[:, None]).long()
        nz = dict(u_noise=torch.rand(B, N, 6, generator=gen).cuda(), u_noise_edges=torch.rand(B, P, 2, generator=gen).cuda(),
                  u_noise_virtual=torch.rand(B, P, 2, generator=gen).cuda())
        for p in params:
            p.grad = None
        z, ldj = model(x.cuda(), adjacency=adj.cuda(), length=length.cuda(), **nz)
        loss_e = loss_fn(z, ldj, length.cuda())
        loss_e.backward()
        ref = [None if p.grad is None else p.grad.clone() for p in params]
        loss_ref = loss_e.detach().clone()
        # autograd graphs of the eager pass keep the parameters' gradient accumulators bound to the stream they were made on
        # (the default stream); none of them may be alive when the step is captured on the capture stream
        del z, ldj, loss_e
        loss_g = step(x.cuda(), adj.cuda(), length.cuda(), **nz)
        assert_close(loss_g, loss_ref, rtol=1e-5, atol=1e-5, what="loss (trial %d)" % trial)
        scale = max(float(r.abs().max()) for r in ref if r is not None)
        for p, r in zip(params, ref):
            if r is None:
                assert p.grad is None or float(p.grad.abs().max()) == 0.0
                continue
            # the scatter reductions of the backward kernels sum in a different order from run to run
            assert_close(p.grad / scale, r / scale, rtol=1e-3, atol=1e-5, what="grad (trial %d)" % trial)
    assert 1 <= step.captures <= 3
