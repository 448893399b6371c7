from .lift_order import (
    aggregate_edge_index,
    aggregate_node_attributes,
    lift_order_edge_index,
    lift_order_edge_index_weighted,
)
from .centrality import (betweenness_centrality, map_to_nodes, path_node_traversals, path_visitation_probabilities,
                         temporal_betweenness_centrality, temporal_closeness_centrality)
from . import centrality, shortest_paths
from .components import connected_components, largest_connected_component
from .rolling_time_window import RollingTimeWindow
from .weisfeiler_leman import WeisfeilerLeman_test
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
    "path_node_traversals",
    "path_visitation_probabilities",
    "map_to_nodes",
    "betweenness_centrality",
    "RollingTimeWindow",
    "centrality",
    "shortest_paths",
    "connected_components",
    "largest_connected_component",
    "WeisfeilerLeman_test",
]
