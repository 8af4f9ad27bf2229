"""Synthetic workloads of BASELINE.json's graph configs, shared by bench.py and tests/ (not part of the product package).

  config 3  graph_coloring        synthetic 3-colour graphs, N = 20, ``GraphNodeFlow`` defaults of
                                  experiments/graph_coloring/train.py:83-88 (8 flows, hidden 384, 4 RGCN attention layers,
                                  K = 8, d = 2), batch 1024 per GPU
  config 4  molecule_generation   ``GraphCNF`` at the Zinc250k shape (N = 38, 9 node types, 3 bond types + none; flows 4,6,6;
                                  hidden 384 / 192; 4 layers; K = 16 / 8; experiments/molecule_generation/train.py:70-77,
                                  graphCNF.py:222-279): log-likelihood pass, GLOBAL batch 512 sharded over the ranks
  config 5  inverse sampling      the same model's reverse pass (graphCNF.py:128-219), GLOBAL batch 8192 sharded over the ranks

Each model can be built twice: from the drop-in classes on the GPU, and from the UNMODIFIED reference in baseline/_ref on the
CPU with the drop-in model's state dict loaded by name (strict) - the CPU baseline and the parity check of the bench.  The
2020 reference needs two runtime patches to run on torch 2.x (the same two tests/golden/make_golden.py
applies; neither changes what is computed): integer division restored in the sparse Edge-GNN index formula
(layers/networks/graph_layers.py:527,668 ``/ 2`` on LongTensors), float64 class-prior bias of the virtual-edge decoder cast
to float32.  No file of the checkout is edited.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types

import numpy as np
import torch

import workload as W

GC = dict(B=1024, N=20, flows=8, hidden=384, layers=4, K=8, D=2, init_batch=256)
MOL = dict(B_fwd=512, B_inv=8192, N=38, node_types=9, edge_types=3, flows="4,6,6", hidden_nodes=384, hidden_edges=192, layers=4,
           K_nodes=16, K_edges=8, D_nodes=6, D_edges=2, init_batch=64, inv_chunk=1024)


class GCDataset:
    @staticmethod
    def num_node_types():
        return 3


class Zinc:
    max_num_nodes = staticmethod(lambda: MOL["N"])
    num_node_types = staticmethod(lambda: MOL["node_types"])
    num_edge_types = staticmethod(lambda: MOL["edge_types"])
    num_max_neighbours = staticmethod(lambda: 4)
    get_node_prior = staticmethod(lambda data_root="data/": np.zeros(MOL["node_types"], dtype=np.float32))
    get_edge_prior = staticmethod(lambda data_root="data/": np.zeros(MOL["edge_types"], dtype=np.float32))


def _enc(d):
    return {"use_dequantization": False, "use_variational": False, "use_decoder": False, "num_dimensions": d,
            "flow_config": {"num_flows": 0, "hidden_layers": 2, "hidden_size": 128}, "decoder_config": {"num_layers": 1, "hidden_size": 64}}


def gc_params():
    return {"categ_encoding": _enc(GC["D"]), "coupling_num_flows": GC["flows"], "coupling_hidden_size": GC["hidden"],
            "coupling_hidden_layers": GC["layers"], "coupling_num_mixtures": GC["K"], "coupling_mask_ratio": 0.5, "coupling_dropout": 0.0}


def mol_params():
    return {"categ_encoding_nodes": _enc(MOL["D_nodes"]), "categ_encoding_edges": _enc(MOL["D_edges"]),
            "coupling_hidden_size_nodes": MOL["hidden_nodes"], "coupling_hidden_size_edges": MOL["hidden_edges"],
            "coupling_num_flows": MOL["flows"], "coupling_hidden_layers": MOL["layers"], "coupling_num_mixtures_nodes": MOL["K_nodes"],
            "coupling_num_mixtures_edges": MOL["K_edges"], "coupling_mask_ratio": 0.5, "coupling_dropout": 0.0,
            "encoding_virtual_num_flows": 0}


def gc_graphs(gen, B, N=GC["N"], p=0.15):
    """(x [B,N] colours, adjacency [B,N,N] in {0,1}, length [B]): random graphs of 11..N-1 nodes, edge probability p."""
    length = torch.randint(11, N, (B,), generator=gen)
    up = torch.triu((torch.rand(B, N, N, generator=gen) < p).long(), diagonal=1)
    adj = up + up.transpose(1, 2)
    valid = torch.arange(N)[None, :] < length[:, None]
    x = torch.randint(0, 3, (B, N), generator=gen)
    return x, adj * (valid[:, :, None] & valid[:, None, :]).long(), length


def molecules(gen, B, N=MOL["N"]):
    """Zinc-shaped synthetic molecules: random spanning tree + ring closures, degree <= 4, 20..N atoms, bond types 1..3.
    Vectorised over the batch (atom i attaches to a uniformly drawn earlier atom with a free valence)."""
    length = torch.randint(20, N + 1, (B,), generator=gen)
    adj = torch.zeros(B, N, N, dtype=torch.long)
    deg = torch.zeros(B, N, dtype=torch.long)
    rows = torch.arange(B)
    for i in range(1, N):
        score = torch.rand(B, i, generator=gen) - (deg[:, :i] >= 4).float() * 2.0       # full atoms are never chosen
        j = score.argmax(dim=1)
        t = torch.randint(1, 4, (B,), generator=gen)
        on = (i < length) & (deg[rows, j] < 4)
        t = t * on.long()
        adj[rows, i, j] = t
        adj[rows, j, i] = t
        deg[rows, i] += on.long()
        deg[rows, j] += on.long()
    for _ in range(3):                                                                 # ring closures
        i = (torch.rand(B, generator=gen) * length).long()
        j = (torch.rand(B, generator=gen) * length).long()
        ok = (i != j) & (adj[rows, i, j] == 0) & (deg[rows, i] < 4) & (deg[rows, j] < 4)
        adj[rows, i, j] += ok.long()
        adj[rows, j, i] += ok.long()
        deg[rows, i] += ok.long()
        deg[rows, j] += ok.long()
    x = torch.randint(0, MOL["node_types"], (B, N), generator=gen) * (torch.arange(N)[None, :] < length[:, None]).long()
    return x, adj, length


def _quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


# --------------------------------------------------------------------------------------------------------------------
# drop-in models on the GPU
# --------------------------------------------------------------------------------------------------------------------
def build_gc_model(device, seed=0):
    from categoricalnf_b200.experiments.graph_coloring import GraphNodeFlow
    torch.manual_seed(seed)
    np.random.seed(seed)
    return _quiet(GraphNodeFlow, gc_params(), GCDataset).to(device).eval()


def build_mol_model(device, seed=0):
    from categoricalnf_b200.experiments.molecule_generation import GraphCNF
    torch.manual_seed(seed)
    np.random.seed(seed)
    return _quiet(GraphCNF, mol_params(), Zinc).to(device).eval()


def data_init(model, x, adj, length):
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        model.initialize_data_dependent([(x, {"adjacency": adj, "length": length})])
    return model


# --------------------------------------------------------------------------------------------------------------------
# the unmodified reference on the CPU
# --------------------------------------------------------------------------------------------------------------------
def _reference_graph_layers():
    """layers/networks/graph_layers.py of baseline/_ref, loaded from its source with ``/ 2`` -> ``// 2`` in the two pair-index
    formulas (LongTensor true division became float in torch 1.6; index_select rejects it)."""
    path = os.path.join(W.REF_ROOT, "layers", "networks", "graph_layers.py")
    src = open(path).read()
    assert src.count("* edge_indices[...,0]) / 2 +") == 2
    src = src.replace("* edge_indices[...,0]) / 2 +", "* edge_indices[...,0]) // 2 +")
    mod = types.ModuleType("layers.networks.graph_layers")
    mod.__file__ = path
    exec(compile(src, path, "exec"), mod.__dict__)
    return mod


def import_reference_graph_models():
    """(GraphNodeFlow, GraphCNF) classes of the reference.  Not in a process where ``categoricalnf_b200.install`` ran."""
    W.import_reference()                                   # sys.path, matplotlib stand-in, sanity checks
    if "layers.networks.graph_layers" not in sys.modules:
        sys.modules["layers.networks.graph_layers"] = _reference_graph_layers()
    from experiments.graph_coloring.graph_node_flow import GraphNodeFlow
    from experiments.molecule_generation.graphCNF import GraphCNF
    assert W.REF_ROOT in os.path.abspath(sys.modules[GraphCNF.__module__].__file__)
    return GraphNodeFlow, GraphCNF


def build_reference_like(model, kind):
    """The reference's own model (CPU, eval) carrying ``model``'s parameters: state dict loaded by name, strict."""
    GraphNodeFlow, GraphCNF = import_reference_graph_models()
    if kind == "gc":
        ref = _quiet(GraphNodeFlow, gc_params(), GCDataset)
    else:
        ref = _quiet(GraphCNF, mol_params(), Zinc)
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    ref.load_state_dict(sd, strict=True)
    if kind == "mol":
        bias = ref.edge_virtual_decoder.layers.main_net[-1].bias
        bias.data = bias.data.float()
    return ref.eval()


class NoiseRecorder:
    """Replaces ``Uniform.sample`` of the reference encodings' priors: draws from a seeded generator and keeps the draws, so
    the drop-in model can be run on identical noise through its ``u_noise*`` arguments."""

    def __init__(self, encodings, seed):
        self.gen = torch.Generator().manual_seed(seed)
        self.draws = []
        for e in encodings:
            e.prior_distribution.distribution.sample = self

    def __call__(self, sample_shape=torch.Size()):
        self.draws.append(torch.rand(sample_shape, generator=self.gen))
        return self.draws[-1]


def reference_encodings(ref, kind):
    return [ref.node_embed_flow] if kind == "gc" else [ref.node_encoding, ref.edge_attr_encoding, ref.edge_virtual_encoding]
