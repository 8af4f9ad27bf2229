"""Adjacency matrix <-> node-pair list conversions of GraphCNF (reference experiments/molecule_generation/mutils.py:5-33).

The pair list holds every unordered node pair once, node-major: (0,1), (0,2), ..., (0,N-1), (1,2), ...  The reference
rebuilds the two index tensors from Python lists on every call (O(N^2) list comprehension, :6-7) and fills the adjacency
with a Python loop over the batch (:29-30) - both sit in the middle of the sampling path; here the indices are cached per
(N, device) and the scatter is one vectorised assignment.
"""
import torch

_INDEX_CACHE = {}


def pair_indices(num_nodes, device):
    key = (int(num_nodes), str(device))
    hit = _INDEX_CACHE.get(key)
    if hit is None:
        iu = torch.triu_indices(num_nodes, num_nodes, offset=1, device=device)
        hit = (iu[0].contiguous(), iu[1].contiguous())
        _INDEX_CACHE[key] = hit
    return hit


def get_adjacency_indices(num_nodes, length):
    """-> (mask_valid [B,P] float: both nodes of the pair exist, (x_indices1 [P], x_indices2 [P]))."""
    x1, x2 = pair_indices(num_nodes, length.device)
    mask_valid = ((x1[None, :] < length[:, None]) & (x2[None, :] < length[:, None])).float()
    return mask_valid, (x1, x2)


def adjacency2pairs(adjacency, length):
    """[B,N,N] adjacency -> (edge_pairs [B,P], (x_indices1, x_indices2), mask_valid [B,P])."""
    num_nodes = adjacency.shape[1]
    mask_valid, (x1, x2) = get_adjacency_indices(num_nodes, length)
    edge_pairs = adjacency.reshape(adjacency.shape[0], num_nodes * num_nodes).index_select(1, x1 + x2 * num_nodes)
    return edge_pairs, (x1, x2), mask_valid


def pairs2adjacency(num_nodes, pairs, length, x_indices):
    """[B,P] pair values -> symmetric [B,N,N] int64 adjacency with a zero diagonal."""
    x1, x2 = x_indices
    adjacency = pairs.new_zeros(pairs.size(0), num_nodes, num_nodes)
    adjacency[:, x1, x2] = pairs
    adjacency[:, x2, x1] = pairs
    return adjacency.long()
