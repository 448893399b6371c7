from .data import Data, EdgeIndex
from .graph import Graph
from .index_map import HigherOrderIndexMap, IndexMap
from .multi_order_model import MultiOrderModel
from .path_data import PathData
from .temporal_graph import TemporalGraph

__all__ = ["Data", "EdgeIndex", "Graph", "IndexMap", "HigherOrderIndexMap", "MultiOrderModel", "PathData", "TemporalGraph"]
