"""pathpyg_b200 -- B200-native (sm_100a) replacement of pathpyG's temporal-path -> higher-order
De Bruijn lift -> DBGNN hot path, behind pathpyG's own Python operator surface.

    import pathpyg_b200 as pp
    g = pp.TemporalGraph.from_edge_list([...])
    m = pp.MultiOrderModel.from_temporal_graph(g, delta=5, max_order=2)
    data = m.to_dbgnn_data(max_order=2)
    out = pp.nn.DBGNN(...)(data)
"""
from . import algorithms, io, nn, utils
from .core import Data, EdgeIndex, Graph, HigherOrderIndexMap, IndexMap, MultiOrderModel, PathData, TemporalGraph

__version__ = "0.1.0"
__all__ = ["algorithms", "io", "nn", "utils", "Data", "EdgeIndex", "Graph", "IndexMap", "HigherOrderIndexMap",
           "MultiOrderModel", "PathData", "TemporalGraph"]
