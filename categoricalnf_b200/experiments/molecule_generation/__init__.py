from .graphCNF import GraphCNF
from .graphed import GraphedLogLikelihood
from .mutils import adjacency2pairs, get_adjacency_indices, pairs2adjacency

__all__ = ["GraphCNF", "GraphedLogLikelihood", "adjacency2pairs", "get_adjacency_indices", "pairs2adjacency"]
