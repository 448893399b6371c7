from .dbgnn import DBGNN, BipartiteGraphOperator, GCNConv

__all__ = ["DBGNN", "BipartiteGraphOperator", "GCNConv"]
