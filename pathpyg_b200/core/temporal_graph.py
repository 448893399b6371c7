"""``TemporalGraph`` -- the input container of the temporal lift
(``src/pathpyG/core/temporal_graph.py:33-128``): ``data.edge_index`` / ``data.time`` sorted by time,
``mapping``, ``n``, ``m``.  The two per-edge Python dictionaries of the reference constructor
(``:71-75``) are built on first access instead.
"""
from __future__ import annotations

import numpy as np
import torch

from .data import Data, EdgeIndex
from .graph import Graph
from .index_map import IndexMap


class TemporalGraph(Graph):
    def __init__(self, data: Data, mapping: IndexMap | None = None) -> None:
        self.data = data
        if not isinstance(data.edge_index, EdgeIndex):
            data.edge_index = EdgeIndex(data.edge_index.contiguous(), sparse_size=(data.num_nodes, data.num_nodes))
        # reorder by time (temporal_graph.py:58-63; the reference's argsort is not stable, so any
        # tie order is a valid instance -- a stable one is used here).  On the GPU the order comes from the
        # library's radix sort over the significant bits of the time range.
        t = data.time
        if t.numel() > 1 and not bool((t[1:] >= t[:-1]).all()):
            if t.is_cuda and t.dtype in (torch.int64, torch.float64):
                from .. import ops

                order = ops.stable_argsort(t)
            else:
                order = torch.sort(t, stable=True).indices
            for attr in set(data.edge_attrs()).union({"time"}):
                if attr == "edge_index":
                    data.edge_index = EdgeIndex(data.edge_index.as_tensor()[:, order].contiguous(),
                                                sparse_size=data.edge_index.sparse_size)
                else:
                    data[attr] = data[attr][order]
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

    def __str__(self) -> str:
        return f"Temporal Graph with {self.n} nodes and {self.m} time-stamped events"
