"""CPU restatement of the ingest step in front of the lift path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Reference (relative to ``/root/reference``):
* ``src/pathpyG/io/pandas.py:28-57``     _parse_timestamp
* ``src/pathpyG/io/pandas.py:318-396``   df_to_temporal_graph
* ``src/pathpyG/io/pandas.py:572-599``   read_csv_path_data
* ``src/pathpyG/core/index_map.py:340-372`` IndexMap.to_idxs (one dictionary look-up per id)
* ``src/pathpyG/core/temporal_graph.py:58-63`` the time ordering of TemporalGraph.__init__
* ``src/pathpyG/core/path_data.py:126-159``  PathData.append_walks (via ``oracle.mom.append_walks``)

The reference orders events with ``torch.argsort`` (not stable): the order of events with EQUAL time stamps is
unspecified there.  This restatement (and the fixtures made by ``oracle/ref_loader.io_module``) use the stable order.
Pinned by the known answers of ``tests/io/test_pandas.py:300-349`` and by ``tests/golden/ingest_golden.npz``.
"""
from __future__ import annotations

import ast
import csv

import numpy as np
import pandas as pd
import torch

from . import mom


def parse_timestamp(df: pd.DataFrame, timestamp_format: str = "%Y-%m-%d %H:%M:%S", time_rescale: int = 1) -> None:
    """io/pandas.py:41-57."""
    if pd.api.types.is_string_dtype(df["t"]):
        df["t"] = pd.to_datetime(df["t"], format=timestamp_format)
        df["t"] = df["t"].astype("int64") // time_rescale
        df["t"] = df["t"] - df["t"].min()
    elif df["t"].dtype == "int64" or df["t"].dtype == "float64":
        df["t"] = df["t"] // time_rescale
    elif pd.api.types.is_datetime64_any_dtype(df["t"]):
        df["t"] = df["t"].astype("int64") // time_rescale
        df["t"] = df["t"] - df["t"].min()
    else:
        raise ValueError(f"Column `t` must be of type `object`, `int64`, `float64`, or a datetime type. Found {df['t'].dtype} instead.")


def df_to_temporal_graph(df: pd.DataFrame, multiedges: bool = False, timestamp_format="%Y-%m-%d %H:%M:%S", time_rescale=1,
                         num_nodes=None) -> dict:
    """io/pandas.py:357-396 + temporal_graph.py:58-63.  Returns the fields of the resulting graph."""
    if all(isinstance(x, int) for x in df.columns.values.tolist()):
        df.columns = ["v", "w", "t"] + [f"edge_attr_{i - 2}" for i in range(3, len(df.columns))]     # :360-366
    parse_timestamp(df, timestamp_format, time_rescale)                                             # :369
    if not multiedges:
        df = df.drop_duplicates(subset=["v", "w", "t"])                                             # :373
    node_ids = np.unique(df[["v", "w"]].values)                                                     # :376
    lut = {v: i for i, v in enumerate(node_ids.tolist())}
    edge_index = torch.tensor([[lut[v] for v in row] for row in df[["v", "w"]].values.T.tolist()])  # :378, index_map.py:368
    time = torch.tensor(df["t"].values)                                                             # :379
    order = torch.sort(time, stable=True).indices                                                   # temporal_graph.py:58
    out = dict(node_ids=node_ids, edge_index=edge_index[:, order], time=time[order],
               num_nodes=num_nodes if num_nodes is not None else node_ids.shape[0])
    for col in [c for c in df.columns if c not in ("v", "w", "t")]:                                 # :384-391 (numeric columns)
        out[col if col.startswith("edge_") else "edge_" + col] = torch.tensor(df[col].values)[order]
    return out


def read_csv_path_data(path: str, weight: bool = True, sep: str = ","):
    """io/pandas.py:581-599.  Returns (sorted node ids, oracle.mom.Walks)."""
    with open(path, "r") as f:
        reader = csv.reader(f, delimiter=sep)
        if weight:
            rows = [(row[:-1], ast.literal_eval(row[-1])) for row in reader]
            paths, weights = zip(*rows)
        else:
            paths = list(reader)
            weights = [1.0] * len(paths)
    node_ids = np.unique(np.hstack(paths))                                                          # :591-592
    lut = {v: i for i, v in enumerate(node_ids.tolist())}
    return node_ids, mom.append_walks([[lut[v] for v in p] for p in paths], list(weights))        # :596-598
