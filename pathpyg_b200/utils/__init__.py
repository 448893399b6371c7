from .convert import to_numpy
from .dbgnn import generate_bipartite_edge_index

__all__ = ["generate_bipartite_edge_index", "to_numpy"]
