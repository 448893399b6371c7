"""The step in front of the lift: ``df_to_temporal_graph`` (reference
``src/pathpyG/io/pandas.py:318-396``), ``read_csv_temporal_graph`` (``:511-545``), ``read_csv_path_data``
(``:572-599``), the column parsers they share (``:28-109``) and ``temporal_graph_to_df`` (``:437-471``).

Same results as the reference, different mechanics:
* node ids are factorised ONCE with ``np.unique(return_inverse=True)`` instead of one dictionary look-up per
  end point (``IndexMap.to_idxs``, reference ``core/index_map.py:368``, 1-2 us per id);
* with ``device=`` the index tensors go straight to the GPU and ``TemporalGraph`` puts them in time order there
  with the library's radix sort (``ops.stable_argsort``) -- the reference sorts on the host with ``torch.argsort``;
* n-gram files are parsed in one pass and handed to ``PathData.append_index_walks`` as one flat tensor.
"""
from __future__ import annotations

import ast
import csv
import logging
import re
from typing import Any

import numpy as np
import pandas as pd
import torch

from ..core.data import Data
from ..core.graph import Graph
from ..core.index_map import IndexMap
from ..core.path_data import PathData
from ..core.temporal_graph import TemporalGraph

logger = logging.getLogger("root")

# the reference's column classifiers (io/pandas.py:22-25)
_iterable_re = re.compile(r"^\s*[\[\(].*[\]\)]\s*$")
_number_re = re.compile(r"^\s*[+-]?(\d+(\.\d*)?|\.\d+)([eE][+-]?\d+)?\s*$")
_integer_re = re.compile(r"^\s*[+-]?\d+\s*$")


def _parse_timestamp(df: pd.DataFrame, timestamp_format: str = "%Y-%m-%d %H:%M:%S", time_rescale: int = 1) -> None:
    """io/pandas.py:28-57 -- in place on column ``t``."""
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
        raise ValueError("Column `t` must be of type `object`, `int64`, `float64`, or a datetime type. "
                         f"Found {df['t'].dtype} instead.")


def _parse_df_column(df: pd.DataFrame, data: Data, attr: str, idx=None, prefix: str = "") -> None:
    """io/pandas.py:60-109 -- one DataFrame column -> ``data[prefix + attr]`` (tensor, or numpy array for text)."""
    if idx is None:
        idx = np.arange(len(df))
    dev = data.edge_index.device
    col = df[attr]
    if col.dtype == "object" or pd.api.types.is_string_dtype(col):
        first = col.values[0]
        if isinstance(first, str):
            if _iterable_re.match(str(first)):
                data[prefix + attr] = torch.tensor([ast.literal_eval(x) for x in col.values[idx]], device=dev)
            elif _number_re.match(str(first)):
                kind = int if _integer_re.match(str(first)) else float
                data[prefix + attr] = torch.tensor(col.values.astype(kind)[idx], device=dev)
            else:
                data[prefix + attr] = np.array(col.values.astype(str)[idx])
        elif isinstance(first, (list, tuple)):
            data[prefix + attr] = torch.tensor(np.array([np.array(x) for x in col.values[idx]]))
        else:
            raise ValueError(f"Unsupported data type for attribute '{attr}': {type(first)}")
    else:
        data[prefix + attr] = torch.tensor(col.values[idx], device=dev)


def df_to_graph(df: pd.DataFrame, is_undirected: bool = False, multiedges: bool = False, num_nodes: int | None = None,
                device=None) -> Graph:
    """io/pandas.py:109-180: columns ``v``, ``w`` (or the first two of a header-less frame) are the edges, every other
    column an edge attribute.  ``device`` (extension): where the index tensors are created."""
    if all(isinstance(x, int) for x in df.columns.values.tolist()):
        df.columns = ["v", "w"] + [f"edge_attr_{i - 2}" for i in range(2, len(df.columns))]
    if not multiedges and df[["v", "w"]].duplicated().any():
        logger.debug("Data frame contains multiple edges, but multiedges is set to False. Removing duplicates.")
        df = df.drop_duplicates(subset=["v", "w"])
    # sorted distinct ids like IndexMap(np.unique(...)) (:154); the inverse IS mapping.to_idxs of the end points
    node_ids, inverse = np.unique(df[["v", "w"]].values, return_inverse=True)
    mapping = IndexMap(node_ids)
    edge_index = torch.from_numpy(np.ascontiguousarray(inverse.reshape(-1, 2).T.astype(np.int64)))
    if device is not None:
        edge_index = edge_index.to(device)
    data = Data(edge_index=edge_index, num_nodes=num_nodes if num_nodes is not None else int(node_ids.shape[0]))
    for col in [c for c in df.columns if c not in ("v", "w")]:
        _parse_df_column(df=df, data=data, attr=col, prefix="" if col.startswith("edge_") else "edge_")
    g = Graph(data=data, mapping=mapping)
    return g.to_undirected() if is_undirected else g


def add_node_attributes(df: pd.DataFrame, g: Graph) -> None:
    """io/pandas.py:183-234: rows are nodes named in column ``v`` (ids) or ``index`` (indices); every other column
    becomes ``node_<column>``, stored in node order."""
    if "v" in df:
        named = list(df["v"])
    elif "index" in df:
        named = list(df["index"])
    else:
        raise ValueError("DataFrame must either have `index` or `v` column")
    if len(set(named)) < len(named):
        raise ValueError("DataFrame cannot contain multiple attribute values for single node")
    if set(named) != (set(g.nodes) if "v" in df else set(range(g.n))):
        raise ValueError("Mismatch between nodes in DataFrame and nodes in graph")
    node_idx = g.mapping.to_idxs(named).tolist() if "v" in df else named
    for attr in [c for c in df.columns if c not in ("v", "index")]:
        _parse_df_column(df=df, data=g.data, idx=node_idx, attr=attr, prefix="" if attr.startswith("node_") else "node_")


def add_edge_attributes(df: pd.DataFrame, g: Graph, time_attr: str | None = None) -> None:
    """io/pandas.py:237-315: rows are edges ``v``, ``w`` (plus the time stamp column ``time_attr`` for a temporal
    graph); every other column becomes ``edge_<column>``, taken in the order the rows are listed."""
    if "v" not in df or "w" not in df:
        raise ValueError("Data frame must have columns `v` and `w` for source and target nodes")
    node_ids = set(df["v"]).union(set(df["w"]))
    known = set(g.nodes)
    if not node_ids.issubset(known):
        raise ValueError(f"DataFrame contains nodes {node_ids - known} that do not exist in the graph. "
                         "Please ensure all nodes in the DataFrame are present in the graph.")
    if g.m != len(df):
        raise ValueError(f"DataFrame contains {len(df)} edges, but the graph has {g.m} edges. "
                         "Please ensure the DataFrame matches the number of edges in the graph.")
    src = g.mapping.to_idxs(df["v"].tolist()).tolist()
    tgt = g.mapping.to_idxs(df["w"].tolist()).tolist()
    edge_attrs = [c for c in df.columns if c not in ("v", "w")]
    if time_attr is not None:
        if time_attr not in df:
            raise ValueError(f"Data frame must have column {time_attr} for time stamps")
        edge_attrs.remove(time_attr)
        lut, keys = g.tedge_to_index, list(zip(src, tgt, df[time_attr].values.tolist()))
    else:
        lut, keys = g.edge_to_index, list(zip(src, tgt))
    edge_idx = []
    for key in keys:
        at = lut.get(key)
        if at is None:
            raise ValueError(f"Edge ({key[0]}, {key[1]}) does not exist" + (f" at time {key[2]}" if time_attr else "")
                             + " in the graph.")
        edge_idx.append(at)
    picked = df.iloc[edge_idx]
    for attr in edge_attrs:
        _parse_df_column(df=picked, data=g.data, attr=attr, prefix="" if attr.startswith("edge_") else "edge_")


def df_to_temporal_graph(df: pd.DataFrame, multiedges: bool = False, timestamp_format="%Y-%m-%d %H:%M:%S", time_rescale=1,
                         num_nodes: int | None = None, device=None) -> TemporalGraph:
    """io/pandas.py:318-396.  ``device`` (extension): where the index tensors are created; with a CUDA device the
    time ordering of ``TemporalGraph`` runs there."""
    if all(isinstance(x, int) for x in df.columns.values.tolist()):
        logger.info("Interpreting first three columns as v, w, t")
        df.columns = ["v", "w", "t"] + [f"edge_attr_{i - 2}" for i in range(3, len(df.columns))]
    _parse_timestamp(df=df, timestamp_format=timestamp_format, time_rescale=time_rescale)
    if not multiedges:
        df = df.drop_duplicates(subset=["v", "w", "t"])
    # np.unique sorts the ids like the reference's IndexMap(np.unique(...)) (:376); the inverse IS mapping.to_idxs
    node_ids, inverse = np.unique(df[["v", "w"]].values, return_inverse=True)
    mapping = IndexMap(node_ids)
    edge_index = torch.from_numpy(np.ascontiguousarray(inverse.reshape(-1, 2).T.astype(np.int64)))
    time = torch.tensor(df["t"].values)
    if device is not None:
        edge_index, time = edge_index.to(device), time.to(device)
    data = Data(edge_index=edge_index, time=time, num_nodes=num_nodes if num_nodes is not None else int(node_ids.shape[0]))
    for col in [c for c in df.columns if c not in ("v", "w", "t")]:
        _parse_df_column(df=df, data=data, attr=col, prefix="" if col.startswith("edge_") else "edge_")
    return TemporalGraph(data=data, mapping=mapping)


def read_csv_temporal_graph(filename: str, sep: str = ",", header: bool = True, timestamp_format: str = "%Y-%m-%d %H:%M:%S",
                            time_rescale: int = 1, **kwargs: Any) -> TemporalGraph:
    """io/pandas.py:511-545."""
    df = pd.read_csv(filename, header=0 if header else None, sep=sep)
    return df_to_temporal_graph(df, timestamp_format=timestamp_format, time_rescale=time_rescale, **kwargs)


def _attr_list(val):
    return val.cpu().numpy().tolist() if isinstance(val, torch.Tensor) else val.tolist()


def graph_to_df(graph: Graph, node_indices: bool = False) -> pd.DataFrame:
    """io/pandas.py:399-428: one row per edge, edge attributes as extra columns."""
    ei = graph.data.edge_index.as_tensor().cpu().numpy()
    v, w = (ei[0], ei[1]) if node_indices else (graph.mapping.to_ids(ei[0]), graph.mapping.to_ids(ei[1]))
    return pd.DataFrame({"v": v, "w": w, **{a: _attr_list(graph.data[a]) for a in graph.edge_attrs()}})


def read_csv_graph(filename: str, sep: str = ",", header: bool = True, is_undirected: bool = False,
                   multiedges: bool = False, **kwargs: Any) -> Graph:
    """io/pandas.py:472-508."""
    df = pd.read_csv(filename, header=0 if header else None, sep=sep)
    return df_to_graph(df, is_undirected=is_undirected, multiedges=multiedges, **kwargs)


def write_csv(graph, node_indices: bool = False, path_or_buf: Any = None, **pdargs: Any) -> None:
    """io/pandas.py:548-569: the edge table of a graph or temporal graph as csv."""
    if isinstance(graph, TemporalGraph):
        frame = temporal_graph_to_df(graph=graph, node_indices=node_indices)
    else:
        frame = graph_to_df(graph=graph, node_indices=node_indices)
    frame.to_csv(index=False, path_or_buf=path_or_buf, **pdargs)


def temporal_graph_to_df(graph: TemporalGraph, node_indices: bool = False) -> pd.DataFrame:
    """io/pandas.py:437-471: one row per time-stamped edge, edge attributes as extra columns."""
    ei = graph.data.edge_index.as_tensor().cpu().numpy()
    if node_indices:
        v, w = ei[0], ei[1]
    else:
        v, w = graph.mapping.to_ids(ei[0]), graph.mapping.to_ids(ei[1])
    df = pd.DataFrame({"v": v, "w": w, "t": graph.data.time.cpu().numpy()})
    for attr in graph.edge_attrs():
        df[attr] = _attr_list(graph.data[attr])
    return df


def read_csv_path_data(path_or_buf: Any = None, weight: bool = True, sep=",", device=None) -> PathData:
    """io/pandas.py:572-599: one walk per line, optionally followed by its count."""
    with open(path_or_buf, "r") as f:
        rows = list(csv.reader(f, delimiter=sep))
    if weight:
        paths = [row[:-1] for row in rows]
        weights = [ast.literal_eval(row[-1]) for row in rows]
    else:
        paths, weights = rows, [1.0] * len(rows)
    flat = np.hstack(paths) if paths else np.empty(0, dtype=str)
    node_ids, inverse = np.unique(flat, return_inverse=True)     # sorted ids + index of every occurrence (:591-592)
    mapping = IndexMap()
    mapping.add_ids(node_ids)
    pathdata = PathData(mapping, device)
    if paths:
        lengths = torch.tensor([len(p) for p in paths], device=device)
        pathdata.append_index_walks(torch.from_numpy(inverse.astype(np.int64)).to(device) if device is not None
                                    else torch.from_numpy(inverse.astype(np.int64)), lengths, torch.tensor(weights, device=device))
    return pathdata
