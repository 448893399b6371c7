"""Ingest of the lift path's inputs (reference ``src/pathpyG/io/pandas.py``): time-stamped edge lists and n-gram
path files -> ``TemporalGraph`` / ``PathData`` with the index tensors placed on the GPU."""
from .pandas import df_to_temporal_graph, read_csv_path_data, read_csv_temporal_graph, temporal_graph_to_df

__all__ = ["df_to_temporal_graph", "read_csv_temporal_graph", "read_csv_path_data", "temporal_graph_to_df"]
