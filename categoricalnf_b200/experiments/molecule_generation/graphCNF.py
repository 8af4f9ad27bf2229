"""GraphCNF for molecule generation (BASELINE configs 4 and 5; reference experiments/molecule_generation/graphCNF.py).

Three steps, each a stack of [ActNorm, InvertibleConv, mixture coupling] blocks:
  1. node types -> latents, coupling network = relational GCN over the full typed adjacency;
  2. edge attributes (bond types) -> latents on the pairs that are bonded, node+edge couplings through an Edge-GNN with
     edge-driven sigmoid attention;
  3. "virtual" edges (no bond) -> latents on all remaining pairs, node+edge couplings through an Edge-GNN with
     query-key attention; a small decoder separates real from virtual edges.
``forward`` returns the node latents and the ldj INCLUDING the prior log-probability of the edge latents (:249-251);
``reverse=True`` samples edge latents, decodes which pairs are bonded, their types and finally the node types, returning
``((node types, adjacency), ldj)``.  Constructor, attribute and state-dict names follow the reference.
"""
import numpy as np
import torch
import torch.nn as nn

from ...layers.categorical_encoding.decoder import DecoderLinear
from ...layers.categorical_encoding.linear_encoding import LinearCategoricalEncoding
from ...layers.categorical_encoding.mutils import create_encoding
from ...layers.flows.activation_normalization import ActNormFlow
from ...layers.flows.coupling_layer import CouplingLayer
from ...layers.flows.distributions import create_prior_distribution
from ...layers.flows.flow_layer import FlowLayer
from ...layers.flows.flow_model import FlowModel
from ...layers.flows.mixture_cdf_layer import MixtureCDFCoupling
from ...layers.flows.node_edge_coupling import NodeEdgeCoupling, NodeEdgeFlowWrapper
from ...layers.flows.permutation_layers import InvertibleConv
from ...layers.networks.graph_layers import (Edge2NodeAttnLayer, Edge2NodeQKVAttnLayer, EdgeGNN, EdgeGNNLayer,
                                             Node2EdgePlainLayer, RelationGraphConv, RGCNNet)
from .mutils import adjacency2pairs, get_adjacency_indices, pairs2adjacency


def _param(params, key, default):
    return params[key] if key in params and params[key] is not None else default


def _channel_mask(length, max_len):
    return (torch.arange(max_len, device=length.device).unsqueeze(0) < length.unsqueeze(1)).float().unsqueeze(-1)


class GraphCNF(FlowModel):

    def __init__(self, model_params, dataset_class, **kwargs):
        super().__init__(layers=None, name="GraphCNF")
        self.model_params = model_params
        self.dataset_class = dataset_class
        self._create_layers()
        self.print_overview()

    # -- construction (:38-192) ------------------------------------------------------------------------------------
    def _create_layers(self):
        self.max_num_nodes = self.dataset_class.max_num_nodes()
        self.num_node_types = self.dataset_class.num_node_types()
        self.num_edge_types = self.dataset_class.num_edge_types()
        self.num_max_neighbours = self.dataset_class.num_max_neighbours()
        self.prior_distribution = create_prior_distribution(_param(self.model_params, "prior_distribution", dict()))
        self._create_encoding_layers()
        self._create_step_flows()

    def _create_encoding_layers(self):
        self.node_encoding = create_encoding(self.model_params["categ_encoding_nodes"], dataset_class=self.dataset_class,
                                             vocab_size=self.num_node_types,
                                             category_prior=self.dataset_class.get_node_prior(data_root="data/"))
        self.edge_attr_encoding = create_encoding(self.model_params["categ_encoding_edges"], dataset_class=self.dataset_class,
                                                  vocab_size=self.num_edge_types,      # the virtual edge is not a class here
                                                  category_prior=self.dataset_class.get_edge_prior(data_root="data/"))
        self.encoding_dim_nodes = self.node_encoding.D
        self.encoding_dim_edges = self.edge_attr_encoding.D
        # virtual edges: a single logistic in latent space; which pairs are virtual is decided by a learned decoder
        self.edge_virtual_encoding = LinearCategoricalEncoding(
            num_dimensions=self.encoding_dim_edges,
            flow_config={"num_flows": _param(self.model_params, "encoding_virtual_num_flows", 0), "hidden_layers": 2, "hidden_size": 128},
            dataset_class=self.dataset_class, vocab_size=1)
        self.edge_virtual_decoder = DecoderLinear(num_categories=2, embed_dim=self.encoding_dim_edges, hidden_size=128, num_layers=2,
                                                  class_prior_log=np.log(np.array([0.9, 0.1])))   # molecules are sparse graphs

    def _create_step_flows(self):
        p = self.model_params
        hidden_size_nodes = _param(p, "coupling_hidden_size_nodes", 256)
        hidden_size_edges = _param(p, "coupling_hidden_size_edges", 128)
        num_flows = [int(k) for k in str(_param(p, "coupling_num_flows", "4,6,6")).split(",")]
        hidden_layers = _param(p, "coupling_hidden_layers", 4)
        if isinstance(hidden_layers, str):
            hidden_layers = [int(v) for v in hidden_layers.split(",")] if "," in hidden_layers else [int(hidden_layers)] * 3
        else:
            hidden_layers = [hidden_layers] * 3
        num_mixtures_nodes = _param(p, "coupling_num_mixtures_nodes", 16)
        num_mixtures_edges = _param(p, "coupling_num_mixtures_edges", 16)
        mask_ratio = _param(p, "coupling_mask_ratio", 0.5)
        dropout = _param(p, "coupling_dropout", 0.0)
        Dn, De = self.encoding_dim_nodes, self.encoding_dim_edges

        coupling_mask_nodes = CouplingLayer.create_channel_mask(Dn, ratio=mask_ratio)
        step1_model_func = lambda c_out: RGCNNet(c_in=Dn, c_out=c_out, num_edges=self.num_edge_types, num_layers=hidden_layers[0],
                                                 hidden_size=hidden_size_nodes, max_neighbours=self.dataset_class.num_max_neighbours(),
                                                 dp_rate=dropout, rgc_layer_fun=RelationGraphConv)
        step1 = []
        for _ in range(num_flows[0]):
            step1 += [ActNormFlow(Dn), InvertibleConv(Dn),
                      MixtureCDFCoupling(c_in=Dn, mask=coupling_mask_nodes, model_func=step1_model_func, block_type="RelationGraphConv",
                                         num_mixtures=num_mixtures_nodes, regularizer_max=3.5, regularizer_factor=2)]
        self.step1_flows = nn.ModuleList(step1)

        coupling_mask_edges = CouplingLayer.create_channel_mask(De, ratio=mask_ratio)

        def edge2node_layer_func(step_idx):
            if step_idx == 1:
                return lambda: Edge2NodeAttnLayer(hidden_size_nodes=hidden_size_nodes, hidden_size_edges=hidden_size_edges, skip_config=2)
            return lambda: Edge2NodeQKVAttnLayer(hidden_size_nodes=hidden_size_nodes, hidden_size_edges=hidden_size_edges, skip_config=2)

        node2edge_layer_func = lambda: Node2EdgePlainLayer(hidden_size_nodes=hidden_size_nodes, hidden_size_edges=hidden_size_edges,
                                                           skip_config=2)

        def get_model_func(step_idx):
            return lambda c_out_nodes, c_out_edges: EdgeGNN(
                c_in_nodes=Dn, c_in_edges=De, c_out_nodes=c_out_nodes, c_out_edges=c_out_edges,
                edge_gnn_layer_func=lambda: EdgeGNNLayer(edge2node_layer_func=edge2node_layer_func(step_idx),
                                                         node2edge_layer_func=node2edge_layer_func),
                max_neighbours=self.dataset_class.num_max_neighbours(), num_layers=hidden_layers[step_idx])

        actnorm_layer = lambda: NodeEdgeFlowWrapper(node_flow=ActNormFlow(c_in=Dn), edge_flow=ActNormFlow(c_in=De))
        permut_layer = lambda: NodeEdgeFlowWrapper(node_flow=InvertibleConv(c_in=Dn), edge_flow=InvertibleConv(c_in=De))
        coupling_layer = lambda step_idx: NodeEdgeCoupling(
            c_in_nodes=Dn, c_in_edges=De, mask_nodes=coupling_mask_nodes, mask_edges=coupling_mask_edges,
            num_mixtures_nodes=num_mixtures_nodes, num_mixtures_edges=num_mixtures_edges, model_func=get_model_func(step_idx),
            regularizer_max=3.5, regularizer_factor=2)
        self.step2_flows = nn.ModuleList([m for _ in range(num_flows[1]) for m in (actnorm_layer(), permut_layer(), coupling_layer(1))])
        self.step3_flows = nn.ModuleList([m for _ in range(num_flows[2]) for m in (actnorm_layer(), permut_layer(), coupling_layer(2))])

    # -- execution (:198-353) --------------------------------------------------------------------------------------
    def _run_layer(self, layer, z, reverse, ldj, ldj_per_layer=None, **kwargs):
        res = layer(z, reverse=reverse, **kwargs)
        z, layer_ldj = res[0], res[1]
        if ldj_per_layer is not None:
            ldj_per_layer.append(res[2] if len(res) == 3 else layer_ldj)
        return z, ldj + layer_ldj

    def _run_node_edge_layer(self, layer, z_nodes, z_edges, reverse, ldj, ldj_per_layer=None, **kwargs):
        res = layer(z_nodes=z_nodes, z_edges=z_edges, reverse=reverse, **kwargs)
        if ldj_per_layer is not None:
            ldj_per_layer.append(res[3] if len(res) == 4 else res[2])
        return res[0], res[1], ldj + res[2]

    def forward(self, z, adjacency=None, ldj=None, reverse=False, get_ldj_per_layer=False, length=None, sample_temp=1.0,
                z_edges_init=None, **kwargs):
        """``z_edges_init`` (reverse only, optional): edge latents to use instead of a fresh prior sample - the hook that
        lets parity tests replay the reference's draw."""
        z_nodes = z
        if ldj is None:
            ldj = z_nodes.new_zeros(z_nodes.size(0), dtype=torch.float32)
        if length is not None:
            kwargs["length"] = length
            kwargs["channel_padding_mask"] = _channel_mask(length, z_nodes.size(1))
        ldj_per_layer = []
        if not reverse:
            z_nodes, ldj = self._step1_forward(z_nodes, adjacency, ldj, False, ldj_per_layer, **kwargs)
            pre = kwargs.pop("cnf_edge_masks", None)     # precomputed by graphed.GraphedLogLikelihood (static buffers)
            if pre is None:
                z_edges_disc, x_indices, mask_valid = adjacency2pairs(adjacency=adjacency, length=length)
                mask_bond = mask_valid * (z_edges_disc != 0).to(mask_valid.dtype)
            else:
                z_edges_disc, x_indices, mask_valid, mask_bond = pre
            kwargs["mask_valid"] = mask_bond
            kwargs["x_indices"] = x_indices
            binary_adjacency = (adjacency > 0).long()
            z_nodes, z_edges, ldj = self._step2_forward(z_nodes, z_edges_disc, ldj, False, ldj_per_layer,
                                                        binary_adjacency=binary_adjacency, **kwargs)
            kwargs["mask_valid"] = mask_valid
            virtual_edge_mask = mask_valid * (z_edges_disc == 0).float()
            z_nodes, z_edges, ldj = self._step3_forward(z_nodes, z_edges, ldj, False, ldj_per_layer, virtual_edge_mask, **kwargs)
            adjacency_log_prob = (self.prior_distribution.log_prob(z_edges) * mask_valid.unsqueeze(dim=-1)).sum(dim=[1, 2])
            ldj = ldj + adjacency_log_prob
            ldj_per_layer.append({"adjacency_log_prob": adjacency_log_prob})
        else:
            batch_size, num_nodes = z_nodes.size(0), z_nodes.size(1)
            mask_valid, x_indices = get_adjacency_indices(num_nodes=num_nodes, length=length)
            kwargs["mask_valid"] = mask_valid
            kwargs["x_indices"] = x_indices
            if z_edges_init is not None:
                z_edges = z_edges_init
            else:
                z_edges = self.prior_distribution.sample(shape=(batch_size, mask_valid.size(1), self.encoding_dim_edges),
                                                         temp=sample_temp).to(z.device)
            z_nodes, z_edges, ldj, mask_valid = self._step3_forward(z_nodes, z_edges, ldj, True, ldj_per_layer, **kwargs)
            binary_adjacency = pairs2adjacency(num_nodes=num_nodes, pairs=mask_valid, length=length, x_indices=x_indices)
            kwargs["mask_valid"] = mask_valid
            z_nodes, z_edges, ldj = self._step2_forward(z_nodes, z_edges, ldj, True, ldj_per_layer,
                                                        binary_adjacency=binary_adjacency, **kwargs)
            adjacency = pairs2adjacency(num_nodes=num_nodes, pairs=z_edges, length=length, x_indices=x_indices)
            z_nodes, ldj = self._step1_forward(z_nodes, adjacency, ldj, reverse=True, ldj_per_layer=ldj_per_layer, **kwargs)
            z_nodes = (z_nodes, adjacency)
        if get_ldj_per_layer:
            return z_nodes, ldj, ldj_per_layer
        return z_nodes, ldj

    def _step1_forward(self, z_nodes, adjacency, ldj, reverse, ldj_per_layer, **kwargs):
        if not reverse:
            z_nodes, ldj = self._run_layer(self.node_encoding, z_nodes, reverse, ldj=ldj, ldj_per_layer=ldj_per_layer, **kwargs)
            for flow in self.step1_flows:
                z_nodes, ldj = self._run_layer(flow, z_nodes, reverse, ldj=ldj, ldj_per_layer=ldj_per_layer, adjacency=adjacency, **kwargs)
        else:
            for flow in reversed(self.step1_flows):
                z_nodes, ldj = self._run_layer(flow, z_nodes, reverse, ldj=ldj, ldj_per_layer=ldj_per_layer, adjacency=adjacency, **kwargs)
            z_nodes, ldj = self._run_layer(self.node_encoding, z_nodes, reverse, ldj=ldj, ldj_per_layer=ldj_per_layer, **kwargs)
        return z_nodes, ldj

    def _step2_forward(self, z_nodes, z_edges, ldj, reverse, ldj_per_layer, **kwargs):
        kwargs_edge_embed = dict(kwargs, channel_padding_mask=kwargs["mask_valid"].unsqueeze(dim=-1))
        kwargs_edge_embed.pop("u_noise", None)
        if not reverse:
            z_attr = (z_edges - 1).clamp(min=0)
            if "u_noise_edges" in kwargs:
                kwargs_edge_embed["u_noise"] = kwargs["u_noise_edges"]
            z_edges, ldj = self._run_layer(self.edge_attr_encoding, z_attr, reverse, ldj, ldj_per_layer, **kwargs_edge_embed)
            for flow in self.step2_flows:
                z_nodes, z_edges, ldj = self._run_node_edge_layer(flow, z_nodes, z_edges, reverse, ldj, ldj_per_layer, **kwargs)
        else:
            for flow in reversed(self.step2_flows):
                z_nodes, z_edges, ldj = self._run_node_edge_layer(flow, z_nodes, z_edges, reverse, ldj, ldj_per_layer, **kwargs)
            z_edges, ldj = self._run_layer(self.edge_attr_encoding, z_edges, reverse, ldj, ldj_per_layer, **kwargs_edge_embed)
            z_edges = (z_edges + 1) * kwargs["mask_valid"].long()       # pairs that are not bonded -> 0 = no edge
        return z_nodes, z_edges, ldj

    def _step3_forward(self, z_nodes, z_edges, ldj, reverse, ldj_per_layer, virtual_edge_mask=None, **kwargs):
        if not reverse:
            kwargs_no_edge_embed = dict(kwargs, channel_padding_mask=virtual_edge_mask.unsqueeze(dim=-1))
            kwargs_no_edge_embed.pop("u_noise", None)
            if "u_noise_virtual" in kwargs:
                kwargs_no_edge_embed["u_noise"] = kwargs["u_noise_virtual"]
            virt_edges = z_edges.new_zeros(z_edges.shape[:-1], dtype=torch.long)
            z_virtual_edges, ldj = self._run_layer(self.edge_virtual_encoding, virt_edges, reverse, ldj, ldj_per_layer, **kwargs_no_edge_embed)
            z_edges = torch.where(virtual_edge_mask.unsqueeze(dim=-1) == 1, z_virtual_edges, z_edges)
            edge_log_probs = self.edge_virtual_decoder(z_edges)
            edge_ldj = torch.where(virtual_edge_mask == 1, edge_log_probs[..., 0], edge_log_probs[..., 1] * kwargs["mask_valid"]).sum(dim=-1)
            ldj = ldj + edge_ldj * (kwargs["beta"] if "beta" in kwargs else 1.0)
            with torch.no_grad():
                ldj_per_layer.append({"virtual_edges_bpd": np.log2(np.exp(1)) * edge_ldj / kwargs["mask_valid"].sum(dim=-1)})
            for flow in self.step3_flows:
                z_nodes, z_edges, ldj = self._run_node_edge_layer(flow, z_nodes, z_edges, reverse, ldj, ldj_per_layer, **kwargs)
            return z_nodes, z_edges, ldj
        for flow in reversed(self.step3_flows):
            z_nodes, z_edges, ldj = self._run_node_edge_layer(flow, z_nodes, z_edges, reverse, ldj, ldj_per_layer, **kwargs)
        is_edge = self.edge_virtual_decoder(z_edges).argmax(dim=-1)
        mask_valid = kwargs["mask_valid"] * (is_edge == 1).float()
        return z_nodes, z_edges, ldj, mask_valid

    # -- data-dependent initialisation (:356-415) ------------------------------------------------------------------
    def initialize_data_dependent(self, batch_list):
        with torch.no_grad():
            for batch, kwargs in batch_list:
                kwargs["channel_padding_mask"] = _channel_mask(kwargs["length"], batch.shape[1])
            for layer in [self.node_encoding] + list(self.step1_flows):
                batch_list = FlowModel.run_data_init_layer(batch_list, layer)
            for i in range(len(batch_list)):
                z_nodes, kwargs = batch_list[i]
                z_adjacency, x_indices, mask_valid = adjacency2pairs(adjacency=kwargs["adjacency"], length=kwargs["length"])
                attr_mask_valid = mask_valid * (z_adjacency != 0).to(mask_valid.dtype)
                z_edges = self.edge_attr_encoding((z_adjacency - 1).clamp(min=0), reverse=False,
                                                  channel_padding_mask=attr_mask_valid.unsqueeze(dim=-1))[0]
                kwargs.update(original_z_adjacency=z_adjacency, binary_adjacency=(kwargs["adjacency"] > 0).long(),
                              original_mask_valid=mask_valid, mask_valid=attr_mask_valid, x_indices=x_indices)
                batch_list[i] = ([z_nodes, z_edges], kwargs)
            for layer in self.step2_flows:
                batch_list = FlowModel.run_data_init_layer(batch_list, layer)
            for i in range(len(batch_list)):
                (z_nodes, z_edges), kwargs = batch_list[i]
                no_edge = kwargs["original_mask_valid"] * (kwargs["original_z_adjacency"] == 0).float()
                z_no_edges = self.edge_virtual_encoding(torch.zeros_like(kwargs["original_z_adjacency"]), reverse=False,
                                                        channel_padding_mask=no_edge.unsqueeze(dim=-1))[0]
                z_edges = z_edges * (1 - no_edge)[..., None] + z_no_edges * no_edge[..., None]
                kwargs["mask_valid"] = kwargs["original_mask_valid"]
                kwargs.pop("binary_adjacency")
                batch_list[i] = ([z_nodes, z_edges], kwargs)
            for layer in self.step3_flows:
                batch_list = FlowModel.run_data_init_layer(batch_list, layer)

    def need_data_init(self):
        return True

    def print_overview(self):
        if not hasattr(self, "step1_flows"):
            return
        lines = ["(1) Node %s" % self.node_encoding.info()]
        idx = 2
        for name, flows, enc in (("Step 1", self.step1_flows, None), ("Step 2", self.step2_flows, ("Edge attribute", self.edge_attr_encoding)),
                                 ("Step 3", self.step3_flows, ("Virtual Edge", self.edge_virtual_encoding))):
            if enc is not None:
                lines.append("(%i) %s %s" % (idx, enc[0], enc[1].info()))
                idx += 1
            for layer in flows:
                lines.append("(%i) [%s] %s" % (idx, name, layer.info().replace("\n", "\n\t      ")))
                idx += 1
        width = max([20] + [len(s) for s in "\n".join(lines).split("\n")])
        print("=" * width + "\nGraphCNF\n" + "-" * width + "\n" + "\n".join(lines) + "\n" + "=" * width)
