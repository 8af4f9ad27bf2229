"""Joint node + edge mixture coupling of GraphCNF (reference
experiments/molecule_generation/graph_node_edge_coupling.py).

``NodeEdgeCoupling``     one Edge-GNN call produces the mixture parameters of both the node latents
                         ``[B,N,Cn]`` and the edge latents ``[B,pairs,Ce]``; each is then transformed by
                         one ``cnf_mixcdf_fwd`` / ``cnf_mixcdf_inv`` launch (the reference runs
                         ``get_mixt_params`` + ``run_with_params`` twice, :112-140).
``NodeEdgeFlowWrapper``  applies a node flow with ``length`` / ``channel_padding_mask`` and an edge flow
                         with ``edge_length = mask_valid.sum(1)`` / ``mask_valid`` (:158-165).

Constructor signatures, buffers (``mask_nodes``, ``mask_edges``) and parameters
(``scaling_factor_nodes/_edges``, ``mixture_scaling_factor_nodes/_edges``, ``nn.*``) are the reference's
(SURVEY App. A).  Unlike ``MixtureCDFCoupling.forward`` this layer *accumulates* into the incoming ldj (:99).
"""
import torch
import torch.nn as nn

from ... import functional as CF
from ._masks import mask_lists
from .flow_layer import FlowLayer


class NodeEdgeCoupling(FlowLayer):

    def __init__(self, c_in_nodes, c_in_edges, mask_nodes, mask_edges, num_mixtures_nodes, num_mixtures_edges,
                 model_func, regularizer_max=-1, regularizer_factor=1, **kwargs):
        super().__init__()
        self.c_in_nodes, self.c_in_edges = c_in_nodes, c_in_edges
        self.num_mixtures_nodes, self.num_mixtures_edges = num_mixtures_nodes, num_mixtures_edges
        self.regularizer_max, self.regularizer_factor = regularizer_max, regularizer_factor
        self.register_buffer("mask_nodes", mask_nodes)
        self.register_buffer("mask_edges", mask_edges)
        self.c_out_nodes = self.c_in_nodes * (2 + 3 * self.num_mixtures_nodes)
        self.c_out_edges = self.c_in_edges * (2 + 3 * self.num_mixtures_edges)
        self.nn = model_func(c_out_nodes=self.c_out_nodes, c_out_edges=self.c_out_edges)
        self.scaling_factor_nodes = nn.Parameter(torch.zeros(self.c_in_nodes))
        self.scaling_factor_edges = nn.Parameter(torch.zeros(self.c_in_edges))
        self.mixture_scaling_factor_nodes = nn.Parameter(torch.zeros(self.c_in_nodes, self.num_mixtures_nodes))
        self.mixture_scaling_factor_edges = nn.Parameter(torch.zeros(self.c_in_edges, self.num_mixtures_edges))

    def forward(self, z_nodes, z_edges, ldj=None, reverse=False, length=None, channel_padding_mask=None,
                mask_valid=None, x_indices=None, binary_adjacency=None, **kwargs):
        if ldj is None:
            ldj = z_nodes.new_zeros(z_nodes.size(0))
        mask_nodes = self.mask_nodes[None, :min(self.mask_nodes.size(0), z_nodes.size(1)), :]
        mask_edges = self.mask_edges[None, :min(self.mask_edges.size(0), z_edges.size(1)), :]
        nn_nodes_out, nn_edges_out = self.nn(z_nodes=mask_nodes * z_nodes, z_edges=mask_edges * z_edges, length=length,
                                             channel_padding_mask=channel_padding_mask, x_indices=x_indices,
                                             mask_valid=mask_valid, binary_adjacency=binary_adjacency)
        # The reference zeroes the network output at padded nodes / invalid pairs (:64,:78) and the latents after
        # the transform (:74,:88); the kernel skips those positions (no parameter read, zero ldj, z_out = 0).
        z_nodes_out, nodes_ldj, nodes_reg = self._run_mixture_layer(
            z_nodes, nn_nodes_out, "mask_nodes", self.num_mixtures_nodes, self.scaling_factor_nodes,
            self.mixture_scaling_factor_nodes, reverse, channel_padding_mask)
        z_edges_out, edges_ldj, edges_reg = self._run_mixture_layer(
            z_edges, nn_edges_out, "mask_edges", self.num_mixtures_edges, self.scaling_factor_edges,
            self.mixture_scaling_factor_edges, reverse, mask_valid)
        ldj = ldj + nodes_ldj + edges_ldj
        detail_out = {"ldj": ldj}
        if nodes_reg is not None:
            detail_out["regularizer_nodes_ldj"] = nodes_reg
        if edges_reg is not None:
            detail_out["regularizer_edges_ldj"] = edges_reg
        return z_nodes_out, z_edges_out, ldj, detail_out

    def _run_mixture_layer(self, orig_z, nn_out, mask, num_mixtures, scaling_factor, mixture_scaling_factor, reverse,
                           channel_padding_mask, **kwargs):
        # `mask` names the buffer; its host form (the kernel's mask lists, truncated like `mask[None, :min(len, S), :]` of
        # the reference, :52-53) is cached per buffer version - no device read per call
        mask_c, mask_s = mask_lists(self, mask, orig_z.size(1))
        z_out, ldj, reg = CF.mixcdf(orig_z, nn_out, num_mixtures, scaling_factor, mixture_scaling_factor, mask_c=mask_c,
                                    mask_s=mask_s, pad=channel_padding_mask, reverse=reverse, reg_max=self.regularizer_max,
                                    reg_factor=self.regularizer_factor, training=self.training)
        return z_out, ldj, (None if reverse else reg)     # no regulariser on the reverse path (:124-136)

    def info(self):
        ratio_n = self.mask_nodes.sum().item() / self.mask_nodes.numel()
        ratio_e = self.mask_edges.sum().item() / self.mask_edges.numel()
        return "Node+Edge Mixture Coupling Layer - Nodes: c_in=%i, num_mixtures=%2i, mask_ratio=%3.2f\n" % (
            self.c_in_nodes, self.num_mixtures_nodes, ratio_n) + \
            "                                   Edges: c_in=%i, num_mixtures=%2i, mask_ratio=%3.2f" % (
            self.c_in_edges, self.num_mixtures_edges, ratio_e)


class NodeEdgeFlowWrapper(FlowLayer):

    def __init__(self, node_flow, edge_flow):
        super().__init__()
        self.node_flow = node_flow
        self.edge_flow = edge_flow

    def forward(self, z_nodes, z_edges, ldj=None, reverse=False, length=None, channel_padding_mask=None, mask_valid=None,
                **kwargs):
        z_nodes, ldj = self.node_flow(z_nodes, ldj=ldj, reverse=reverse, length=length,
                                      channel_padding_mask=channel_padding_mask, **kwargs)
        edge_length = mask_valid.sum(dim=1)
        if mask_valid.dim() == 2:
            mask_valid = mask_valid.unsqueeze(dim=-1)
        z_edges, ldj = self.edge_flow(z_edges, ldj=ldj, reverse=reverse, length=edge_length,
                                      channel_padding_mask=mask_valid, **kwargs)
        return z_nodes, z_edges, ldj

    def need_data_init(self):
        return self.node_flow.need_data_init() or self.edge_flow.need_data_init()

    def data_init_forward(self, z_nodes, z_edges, channel_padding_mask=None, mask_valid=None, **kwargs):
        if self.node_flow.need_data_init():
            self.node_flow.data_init_forward(z_nodes, channel_padding_mask=channel_padding_mask)
        if self.edge_flow.need_data_init():
            if mask_valid.dim() == 2:
                mask_valid = mask_valid.unsqueeze(dim=-1)
            self.edge_flow.data_init_forward(z_edges, channel_padding_mask=mask_valid)

    def info(self):
        return "FlowWrapper - Node layer: %s\n" % self.node_flow.info() + \
            "              Edge layer: %s" % self.edge_flow.info()
