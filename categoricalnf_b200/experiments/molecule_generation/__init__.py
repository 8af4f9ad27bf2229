from .graphCNF import GraphCNF
from .graphed import GraphedLogLikelihood, GraphedTrainingStep
from .mutils import adjacency2pairs, get_adjacency_indices, pairs2adjacency

__all__ = ["GraphCNF", "GraphedLogLikelihood", "GraphedTrainingStep", "adjacency2pairs", "get_adjacency_indices", "pairs2adjacency"]
