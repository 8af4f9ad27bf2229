"""Graph-colouring flow (BASELINE config 3; reference experiments/graph_coloring/graph_node_flow.py:17-128).

Node colours are encoded into ``d``-dimensional latents and transformed by ``coupling_num_flows`` blocks of
[ActNorm, InvertibleConv, MixtureCDFCoupling conditioned through an attention RGCN on the graph], closed by a last
ActNorm.  Same constructor (``model_params`` dictionary, ``dataset_class`` with ``num_node_types()``), attribute and
state-dict names as the reference, so its checkpoints load unchanged.
"""
import torch
import torch.nn as nn

from ...layers.categorical_encoding.mutils import create_encoding
from ...layers.flows.activation_normalization import ActNormFlow
from ...layers.flows.coupling_layer import CouplingLayer
from ...layers.flows.flow_model import FlowModel
from ...layers.flows.mixture_cdf_layer import MixtureCDFCoupling
from ...layers.flows.permutation_layers import InvertibleConv
from ...layers.networks.graph_layers import RGCNNet, RelationGraphAttention


def _param(params, key, default):
    """general/mutils.py get_param_val: dictionary value, default when absent."""
    return params[key] if key in params else default


def length_masks(length, max_len):
    """(src_key_padding_mask [B,N] bool, True = padding; channel_padding_mask [B,N,1] float) as
    general/mutils.py:279-288 builds them."""
    valid = torch.arange(max_len, device=length.device).unsqueeze(0) < length.unsqueeze(1)
    return ~valid, valid.float().unsqueeze(-1)


class GraphNodeFlow(FlowModel):

    def __init__(self, model_params, dataset_class, **kwargs):
        super().__init__(layers=None, name="GraphCNF (node based)")
        self.model_params = model_params
        self.dataset_class = dataset_class
        self._create_layers()
        self.print_overview()

    def _create_layers(self):
        self.num_node_types = self.dataset_class.num_node_types()
        self.node_embed_flow = create_encoding(self.model_params["categ_encoding"], dataset_class=self.dataset_class,
                                               vocab_size=self.num_node_types)
        self.embed_dim = self.node_embed_flow.D
        self.flow_layers = nn.ModuleList([self.node_embed_flow] + self._create_node_flow_layers())

    def _create_node_flow_layers(self):
        p = self.model_params
        num_flows = _param(p, "coupling_num_flows", 8)
        hidden_size = _param(p, "coupling_hidden_size", 384)
        hidden_layers = _param(p, "coupling_hidden_layers", 4)
        num_mixtures = _param(p, "coupling_num_mixtures", 16)
        mask_ratio = _param(p, "coupling_mask_ratio", 0.5)
        dropout = _param(p, "coupling_dropout", 0.0)
        coupling_mask = CouplingLayer.create_channel_mask(self.embed_dim, ratio=mask_ratio)
        model_func = lambda c_out: RGCNNet(c_in=self.embed_dim, c_out=c_out, num_edges=1, num_layers=hidden_layers,
                                           hidden_size=hidden_size, dp_rate=dropout, rgc_layer_fun=RelationGraphAttention)
        layers = []
        for _ in range(num_flows):
            layers += [ActNormFlow(self.embed_dim),
                       InvertibleConv(self.embed_dim),
                       MixtureCDFCoupling(c_in=self.embed_dim, mask=coupling_mask, model_func=model_func,
                                          block_type="GraphAttentionNet", num_mixtures=num_mixtures,
                                          regularizer_max=3.5,      # keeps the transform accurately invertible
                                          regularizer_factor=2)]
        layers += [ActNormFlow(c_in=self.embed_dim)]
        return layers

    def forward(self, z, adjacency, ldj=None, reverse=False, length=None, **kwargs):
        if length is not None:
            kwargs["src_key_padding_mask"], kwargs["channel_padding_mask"] = length_masks(length, z.size(1))
        return super().forward(z, adjacency=adjacency, ldj=ldj, reverse=reverse, length=length, **kwargs)

    def initialize_data_dependent(self, batch_list):
        """``batch_list``: (z, kwargs) tuples, kwargs holding ``adjacency`` and ``length``."""
        with torch.no_grad():
            for batch, kwargs in batch_list:
                kwargs["src_key_padding_mask"], kwargs["channel_padding_mask"] = length_masks(kwargs["length"], batch.shape[1])
            for layer in self.flow_layers:
                batch_list = FlowModel.run_data_init_layer(batch_list, layer)

    def need_data_init(self):
        return True

    def test_reversibility(self, z, adjacency, length):
        """Embed, run the continuous layers forward then backward; True when latents agree to 1e-2 and ldj to 1e-1
        (the reference's thresholds, graph_node_flow.py:178)."""
        kwargs = dict(length=length, adjacency=adjacency)
        kwargs["src_key_padding_mask"], kwargs["channel_padding_mask"] = length_masks(length, z.size(1))
        with torch.no_grad():
            z_embed, ldj_embed = self.node_embed_flow(z, reverse=False, **kwargs)[:2]
            z_cur, ldj = z_embed, ldj_embed
            for flow in list(self.flow_layers)[1:]:
                res = flow(z_cur, reverse=False, **kwargs)
                z_cur, ldj = res[0], ldj + res[1]      # no running ldj passed: every layer returns its own term
            z_back, ldj_back = z_cur, ldj
            for flow in reversed(list(self.flow_layers)[1:]):
                res = flow(z_back, reverse=True, **kwargs)
                z_back, ldj_back = res[0], ldj_back + res[1]
        ok = bool(((z_back - z_embed).abs() > 1e-2).sum() == 0 and ((ldj_back - ldj_embed).abs() > 1e-1).sum() == 0)
        print("Reversibility test passed" if ok else "[!] ERROR: Coupling layer with given adjacency matrix are not reversible.")
        return ok
