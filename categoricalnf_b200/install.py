"""Make an unmodified checkout of the reference run on the sm_100a kernels.

``install(reference_root)`` puts the checkout on ``sys.path`` and registers the drop-in modules of
:mod:`categoricalnf_b200.layers` under the module names the reference's experiments import
(``layers.flows.coupling_layer`` ...).  Because Python consults ``sys.modules`` first, every
``from layers.flows.mixture_cdf_layer import MixtureCDFCoupling`` in ``experiments/*`` and in the
reference's un-replaced modules (baseline models, tasks, ...) then resolves to the CUDA-backed
class; nothing in the checkout is edited.  Call it before importing anything from the checkout.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

# reference module name -> drop-in module (relative to categoricalnf_b200.layers)
REPLACED = {
    "layers.flows.flow_layer": "flows.flow_layer",
    "layers.flows.flow_model": "flows.flow_model",
    "layers.flows.coupling_layer": "flows.coupling_layer",
    "layers.flows.mixture_cdf_layer": "flows.mixture_cdf_layer",
    "layers.flows.autoregressive_coupling": "flows.autoregressive_coupling",
    "layers.flows.activation_normalization": "flows.activation_normalization",
    "layers.flows.permutation_layers": "flows.permutation_layers",
    "layers.flows.distributions": "flows.distributions",
    "layers.flows.sigmoid_layer": "flows.sigmoid_layer",
    "layers.categorical_encoding.decoder": "categorical_encoding.decoder",
    "layers.categorical_encoding.linear_encoding": "categorical_encoding.linear_encoding",
    "layers.categorical_encoding.variational_encoding": "categorical_encoding.variational_encoding",
    "layers.categorical_encoding.variational_dequantization": "categorical_encoding.variational_dequantization",
    "layers.categorical_encoding.mutils": "categorical_encoding.mutils",
    # GraphCNF's joint node+edge coupling lives under experiments/ upstream but is hot-path row a14
    "experiments.molecule_generation.graph_node_edge_coupling": "flows.node_edge_coupling",
    # coupling networks of the graph flows: RGCNNet, RelationGraph*, GNNSkipConnection, EdgeGNN and its layers
    "layers.networks.graph_layers": "networks.graph_layers",
}

# callers rebuilt on the drop-in layers; reference module name -> categoricalnf_b200 module
REPLACED_CALLERS = {
    "experiments.graph_coloring.graph_node_flow": "categoricalnf_b200.experiments.graph_coloring.graph_node_flow",
    "experiments.molecule_generation.graphCNF": "categoricalnf_b200.experiments.molecule_generation.graphCNF",
    "experiments.molecule_generation.mutils": "categoricalnf_b200.experiments.molecule_generation.mutils",
}


def _stub_matplotlib():
    """The reference imports matplotlib in a few callers (general/task.py:8, ...) without using it
    on any path run here; provide an empty stand-in when the real package is absent."""
    try:
        import matplotlib  # noqa: F401
        return False
    except ImportError:
        pass
    mpl = types.ModuleType("matplotlib")
    mpl.use = lambda *a, **k: None
    pyplot = types.ModuleType("matplotlib.pyplot")
    colors = types.ModuleType("matplotlib.colors")
    colors.hsv_to_rgb = lambda x: x
    mpl.pyplot, mpl.colors = pyplot, colors
    sys.modules.update({"matplotlib": mpl, "matplotlib.pyplot": pyplot, "matplotlib.colors": colors})
    return True


def install(reference_root: str, stub_missing: bool = True) -> dict:
    """Register the drop-in modules for the checkout at ``reference_root``.  Returns the mapping
    that was installed.  Raises if one of the reference modules was imported already (its classes
    would be bound in the importers and the patch would be silently partial)."""
    root = os.path.abspath(reference_root)
    if not os.path.isdir(os.path.join(root, "layers", "flows")):
        raise FileNotFoundError("%s does not look like a CategoricalNF checkout (layers/flows missing)" % root)
    from . import _lib
    _lib.load()     # fail now, loudly, if the CUDA library is missing
    if root not in sys.path:
        sys.path.insert(0, root)
    if stub_missing:
        _stub_matplotlib()
    installed = {}
    targets = {ref_name: "categoricalnf_b200.layers." + ours for ref_name, ours in REPLACED.items()}
    targets.update(REPLACED_CALLERS)
    for ref_name, ours in targets.items():
        mod = importlib.import_module(ours)
        prev = sys.modules.get(ref_name)
        if prev is not None and prev is not mod:
            raise RuntimeError("%s was imported before categoricalnf_b200.install(); call install() first" % ref_name)
        sys.modules[ref_name] = mod
        installed[ref_name] = mod.__name__
    return installed


def uninstall() -> None:
    for ref_name in list(REPLACED) + list(REPLACED_CALLERS):
        mod = sys.modules.get(ref_name)
        if mod is not None and mod.__name__.startswith("categoricalnf_b200."):
            del sys.modules[ref_name]
