"""Readers and writers either side of the lift path (reference ``src/pathpyG/io/pandas.py``): edge tables, time-stamped
edge lists and n-gram path files <-> ``Graph`` / ``TemporalGraph`` / ``PathData``, index tensors optionally placed on
the GPU (``device=``)."""
from .pandas import (add_edge_attributes, add_node_attributes, df_to_graph, df_to_temporal_graph, graph_to_df,
                     read_csv_graph, read_csv_path_data, read_csv_temporal_graph, temporal_graph_to_df, write_csv)

__all__ = ["add_edge_attributes", "add_node_attributes", "df_to_graph", "df_to_temporal_graph", "graph_to_df",
           "read_csv_graph", "read_csv_path_data", "read_csv_temporal_graph", "temporal_graph_to_df", "write_csv"]
