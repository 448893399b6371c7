from .lift_order import (
    aggregate_edge_index,
    aggregate_node_attributes,
    lift_order_edge_index,
    lift_order_edge_index_weighted,
)
from .centrality import temporal_betweenness_centrality, temporal_closeness_centrality
from .temporal import lift_order_temporal, temporal_shortest_paths

__all__ = [
    "aggregate_edge_index",
    "aggregate_node_attributes",
    "lift_order_edge_index",
    "lift_order_edge_index_weighted",
    "lift_order_temporal",
    "temporal_shortest_paths",
    "temporal_closeness_centrality",
    "temporal_betweenness_centrality",
]
