from .graph_node_flow import GraphNodeFlow

__all__ = ["GraphNodeFlow"]
