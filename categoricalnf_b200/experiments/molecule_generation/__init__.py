from .graphCNF import GraphCNF
from .mutils import adjacency2pairs, get_adjacency_indices, pairs2adjacency

__all__ = ["GraphCNF", "adjacency2pairs", "get_adjacency_indices", "pairs2adjacency"]
