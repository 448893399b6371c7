"""``TemporalGraph`` -- the input container of the temporal lift
(``src/pathpyG/core/temporal_graph.py:33-128``): ``data.edge_index`` / ``data.time`` sorted by time,
``mapping``, ``n``, ``m``.  The two per-edge Python dictionaries of the reference constructor
(``:71-75``) are built on first access instead.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _staging, ops
from .data import Data, EdgeIndex
from .graph import Graph
from .index_map import IndexMap


class TemporalGraph(Graph):
    def __init__(self, data: Data, mapping: IndexMap | None = None) -> None:
        self.data = data
        if data.num_nodes is None:
            # PyG infers a missing num_nodes from the EdgeIndex's sparse size, else from the largest index
            size = data.edge_index.sparse_size[0] if isinstance(data.edge_index, EdgeIndex) else None
            data.num_nodes = size if size is not None else (int(data.edge_index.max()) + 1 if data.edge_index.numel() else 0)
        if not isinstance(data.edge_index, EdgeIndex):
            data.edge_index = EdgeIndex(data.edge_index.contiguous(), sparse_size=(data.num_nodes, data.num_nodes))
        # reorder by time (temporal_graph.py:58-63; the reference's argsort is not stable, so any
        # tie order is a valid instance -- a stable one is used here).  On the GPU the order comes from the
        # library's radix sort over the significant bits of the time range.
        t = data.time
        if t.numel() > 1 and not bool((t[1:] >= t[:-1]).all()):
            if t.is_cuda and t.dtype in (torch.int64, torch.float64):
                order = ops.stable_argsort(t)
            else:
                order = torch.sort(t, stable=True).indices
            for attr in set(data.edge_attrs()).union({"time"}):
                if attr == "edge_index":
                    data.edge_index = EdgeIndex(data.edge_index.as_tensor()[:, order].contiguous(),
                                                sparse_size=data.edge_index.sparse_size)
                else:
                    val = data[attr]
                    data[attr] = val[order] if isinstance(val, torch.Tensor) else val[order.cpu().numpy()]
        self.mapping = mapping if mapping is not None else IndexMap()
        self._edge_to_index = None
        self._tedge_to_index = None
        self._sorted_token = self._time_token()  # the time tensor this object has verified / put in order

    def _time_token(self):
        t = self.data.time
        return (t.data_ptr(), t._version, t.numel(), str(t.device))

    def time_is_known_sorted(self) -> bool:
        """True if ``data.time`` is still the tensor the constructor checked (no device pass needed again)."""
        return self._sorted_token is not None and self._sorted_token == self._time_token()

    def to(self, device) -> "TemporalGraph":
        known = self.time_is_known_sorted()
        self.data = self.data.to(device)
        self._sorted_token = self._time_token() if known else None
        return self

    @property
    def edge_to_index(self) -> dict:
        if self._edge_to_index is None:
            rows, cols = self.data.edge_index.as_tensor().cpu().tolist()
            self._edge_to_index = {(r, c): i for i, (r, c) in enumerate(zip(rows, cols))}
        return self._edge_to_index

    @property
    def tedge_to_index(self) -> dict:
        if self._tedge_to_index is None:
            rows, cols = self.data.edge_index.as_tensor().cpu().tolist()
            times = self.data.time.cpu().tolist()
            self._tedge_to_index = {(r, c, t): i for i, (r, c, t) in enumerate(zip(rows, cols, times))}
        return self._tedge_to_index

    @staticmethod
    def from_edge_list(edge_list, num_nodes: int | None = None, device=None) -> "TemporalGraph":
        if len(edge_list) == 0:
            return TemporalGraph(Data(edge_index=torch.empty((2, 0), dtype=torch.long, device=device),
                                      time=torch.empty((0,), dtype=torch.long, device=device), num_nodes=num_nodes))
        edge_array = np.array(edge_list)
        if isinstance(edge_list[0][2], int):
            ts = torch.tensor(edge_array[:, 2].astype(np.int_), device=device)
        else:
            ts = torch.tensor(edge_array[:, 2].astype(np.double), device=device)
        index_map = IndexMap(np.unique(edge_array[:, :2]))
        edge_index = index_map.to_idxs(edge_array[:, :2].T, device=device)
        if not num_nodes:
            num_nodes = index_map.num_ids()
        return TemporalGraph(Data(edge_index=edge_index, time=ts, num_nodes=num_nodes), mapping=index_map)

    @staticmethod
    def from_tensors(edge_index: torch.Tensor, time: torch.Tensor, num_nodes: int, mapping: IndexMap | None = None,
                     **edge_attrs) -> "TemporalGraph":
        """Bulk constructor for index data that is already on its device (no per-edge Python work)."""
        return TemporalGraph(Data(edge_index=edge_index, time=time, num_nodes=num_nodes, **edge_attrs), mapping=mapping)

    @property
    def temporal_edges(self):
        rows, cols = self.data.edge_index.as_tensor().cpu().tolist()
        times = self.data.time.cpu().tolist()
        ids = self.mapping
        return [(ids.to_id(r), ids.to_id(c), t) for r, c, t in zip(rows, cols, times)]

    @property
    def order(self) -> int:
        return 1

    @property
    def m(self) -> int:
        return int(self.data.edge_index.size(1))

    @property
    def start_time(self):
        return self.data.time.min().item()

    @property
    def end_time(self):
        return self.data.time.max().item()

    def shuffle_time(self) -> None:
        """Randomly permute the time stamps (temporal_graph.py:187-189); like the reference, the events are NOT put
        back into time order -- build a new ``TemporalGraph`` from ``data`` for that."""
        self.data.time = self.data.time[torch.randperm(len(self.data.time), device=self.data.time.device)]
        self._tedge_to_index = None
        self._sorted_token = None

    # ---- views (temporal_graph.py:191-316) -----------------------------------------------------------------
    def to_static_graph(self, weighted: bool = False, time_window=None) -> Graph:
        """Time-aggregated graph; ``weighted`` merges repeated edges and counts them in ``edge_weight``
        (library coalesce: radix sort + run reduction)."""
        ei = self.data.edge_index.as_tensor()
        if time_window is not None:
            t = self.data.time
            ei = ei[:, ((t >= time_window[0]) & (t < time_window[1])).nonzero().ravel()]
        n = int(ei.max()) + 1  # the reference sizes the static graph by the largest index present (:207)
        if weighted:
            dev, to_host = _staging.compute_device(ei)
            i, w = ops.coalesce(_staging.up(ei, dev), None, n, None, "sum")
            return Graph(Data(edge_index=EdgeIndex(_staging.down(i, to_host), sparse_size=(n, n)),
                              edge_weight=_staging.down(w, to_host)), self.mapping)
        return Graph.from_edge_index(EdgeIndex(ei, sparse_size=(n, n)), self.mapping)

    def to_undirected(self) -> "TemporalGraph":
        """Every event in both directions at the same time; edge attributes are not carried over (:222-246)."""
        ei = self.data.edge_index.as_tensor()
        both = EdgeIndex(torch.cat([ei, ei.flip(0)], dim=1), sparse_size=self.data.edge_index.sparse_size)
        return TemporalGraph(Data(edge_index=both, time=torch.cat([self.data.time, self.data.time])), mapping=self.mapping)

    def _subset(self, pick, pick_np) -> "TemporalGraph":
        ei = self.data.edge_index
        data = Data(edge_index=EdgeIndex(ei.as_tensor()[:, pick], sparse_size=ei.sparse_size), time=self.data.time[pick])
        for attr in self.node_attrs():
            data[attr] = self.data[attr]
        for attr in self.edge_attrs():
            val = self.data[attr]
            data[attr] = val[pick] if isinstance(val, torch.Tensor) else val[pick_np]
        return TemporalGraph(data=data, mapping=self.mapping)

    def get_batch(self, start_idx: int, end_idx: int) -> "TemporalGraph":
        """Events ``start_idx .. end_idx - 1`` of the time-ordered list (:248-283)."""
        return self._subset(slice(start_idx, end_idx), slice(start_idx, end_idx))

    def get_window(self, start_time, end_time) -> "TemporalGraph":
        """Events with ``start_time <= t < end_time`` (:285-316)."""
        mask = (self.data.time >= start_time) & (self.data.time < end_time)
        return self._subset(mask, mask.cpu().numpy())

    def __getitem__(self, key):
        """Node, edge, temporal-edge or graph attribute (:318-342); a 3-tuple addresses the LAST event of an edge."""
        if not isinstance(key, tuple):
            if key in self.data.keys():
                return self.data[key]
            raise KeyError(key + " is not a graph attribute")
        if key[0] in self.node_attrs():
            return self.data[key[0]][self.mapping.to_idx(key[1])]
        if key[0] in self.edge_attrs():
            v, w = self.mapping.to_idx(key[1]), self.mapping.to_idx(key[2])
            if len(key) == 3:
                return self.data[key[0]][self.edge_to_index[v, w]]
            return self.data[key[0]][self.tedge_to_index[v, w, key[3]]]
        raise KeyError(key[0] + " is not a node or edge attribute")

    def __str__(self) -> str:
        from pprint import pformat

        def kind(v):
            return str(torch.Tensor) + " -> " + str(v.size()) if isinstance(v, torch.Tensor) else str(type(v))

        ei = self.data.edge_index.as_tensor()
        unique_edges = int(torch.unique(ei[0] * max(self.n, 1) + ei[1]).numel())
        s = (f"Temporal Graph with {self.data.num_nodes} nodes, {unique_edges} unique edges and "
             f"{ei.size(1)} events in [{self.start_time}, {self.end_time}]\n")
        info: dict = {"Node Attributes": {}, "Edge Attributes": {}, "Graph Attributes": {}}
        node_attrs, edge_attrs = self.node_attrs(), self.edge_attrs()
        for k in self.data.keys():
            if k in node_attrs:
                info["Node Attributes"][k] = kind(self.data[k])
            elif k in edge_attrs:
                info["Edge Attributes"][k] = kind(self.data[k])
            elif k not in ("edge_index", "time", "node_sequence"):
                info["Graph Attributes"][k] = kind(self.data[k])
        return s + pformat(info, indent=4, width=160)
