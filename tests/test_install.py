"""``categoricalnf_b200.install``: the drop-in modules are registered under the names the
reference's experiments import.  CPU only; a fake checkout stands in for the reference (which is absent on the GPU box)."""
import importlib
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_install_registers_dropins_under_reference_names(tmp_path):
    (tmp_path / "layers" / "flows").mkdir(parents=True)
    prog = textwrap.dedent("""
        import sys
        sys.path.insert(0, %r)
        import categoricalnf_b200.install as cnf
        installed = cnf.install(%r)
        from layers.flows.mixture_cdf_layer import MixtureCDFCoupling
        from layers.flows.flow_model import FlowModel
        from layers.networks.graph_layers import RGCNNet, EdgeGNN, RelationGraphAttention
        from layers.categorical_encoding.mutils import create_encoding
        from experiments.molecule_generation.graph_node_edge_coupling import NodeEdgeCoupling, NodeEdgeFlowWrapper
        from experiments.molecule_generation.graphCNF import GraphCNF
        from experiments.molecule_generation.mutils import adjacency2pairs, pairs2adjacency
        from experiments.graph_coloring.graph_node_flow import GraphNodeFlow
        for cls in (MixtureCDFCoupling, FlowModel, RGCNNet, EdgeGNN, NodeEdgeCoupling, GraphCNF, GraphNodeFlow):
            assert cls.__module__.startswith("categoricalnf_b200."), cls
        cnf.uninstall()
        assert "layers.flows.mixture_cdf_layer" not in sys.modules
        print("ok", len(installed))
    """) % (ROOT, str(tmp_path))
    out = subprocess.run([sys.executable, "-c", prog], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.strip().startswith("ok")


def test_pair_helpers_match_definition():
    """adjacency <-> pair list conversions (experiments/molecule_generation/mutils.py:5-33) on the CPU."""
    import torch
    sys.path.insert(0, ROOT)
    M = importlib.import_module("categoricalnf_b200.experiments.molecule_generation.mutils")
    g = torch.Generator().manual_seed(0)
    B, N = 3, 7
    up = torch.triu(torch.randint(0, 4, (B, N, N), generator=g), diagonal=1)
    adj = up + up.transpose(1, 2)
    length = torch.tensor([7, 4, 2])
    pairs, (x1, x2), mask_valid = M.adjacency2pairs(adj, length)
    want1 = torch.tensor([i for i in range(N) for j in range(i + 1, N)])
    want2 = torch.tensor([j for i in range(N) for j in range(i + 1, N)])
    assert torch.equal(x1, want1) and torch.equal(x2, want2)
    assert torch.equal(pairs, adj[:, want1, want2])
    assert torch.equal(mask_valid, ((want1[None] < length[:, None]) & (want2[None] < length[:, None])).float())
    assert torch.equal(M.pairs2adjacency(N, pairs, length, (x1, x2)), adj)
    mv, idx = M.get_adjacency_indices(N, length)
    assert torch.equal(mv, mask_valid) and torch.equal(idx[0], x1)


def test_static_pair_context_padding_semantics():
    """Host logic of the CUDA-graph wrappers (CPU tensors): a StaticPairContext with more rows than valid pairs compacts /
    expands like the exact PairContext on its real rows, its padding rows read zeros, write nowhere and carry no gradient."""
    import torch
    from categoricalnf_b200.layers.networks.graph_layers import PairContext, StaticPairContext
    from categoricalnf_b200.experiments.molecule_generation.mutils import pair_indices
    gen = torch.Generator().manual_seed(0)
    B, N, F = 3, 6, 4
    x_indices = pair_indices(N, "cpu")
    P = N * (N - 1) // 2
    mask = (torch.rand(B, P, generator=gen) < 0.5).float()
    mask[1] = 0
    exact = PairContext(x_indices, mask, N)
    padded = StaticPairContext(x_indices, B, N, exact.R + 5, "cpu")
    assert StaticPairContext.count(mask) == exact.R and padded.load(mask)
    assert not StaticPairContext(x_indices, B, N, exact.R - 1, "cpu").load(mask), "too few rows must be refused"
    assert not padded.load(torch.zeros(B, P)), "no valid pair: nothing to pad with"
    assert padded.load(mask)
    feat = torch.randn(B, P, F, generator=gen, requires_grad=True)
    rows = padded.compact(feat)
    assert rows.shape == (exact.R + 5, F)
    assert torch.equal(rows[:exact.R], exact.compact(feat)) and bool((rows[exact.R:] == 0).all())
    assert torch.equal(padded.rev, exact.rev)
    assert torch.equal(padded.node1[:exact.R], exact.node1) and torch.equal(padded.node2[:exact.R], exact.node2)
    assert bool((padded.flat_indices[exact.R:] == exact.flat_indices[0]).all())       # kernels see a valid pair there
    vals = torch.randn(exact.R + 5, F, generator=gen, requires_grad=True)
    out = padded.expand(vals)
    assert torch.equal(out, exact.expand(vals[:exact.R]))
    (out.sum() + rows.sum()).backward()
    assert bool((vals.grad[exact.R:] == 0).all()) and bool((vals.grad[:exact.R] == 1).all())
    assert torch.equal(feat.grad, mask.unsqueeze(-1).expand(B, P, F))
    padded.attach(mask)
    assert PairContext.of(x_indices, mask, N) is padded
