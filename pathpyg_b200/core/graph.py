"""``Graph`` -- the output container of the lift path, with the reference's surface
(``src/pathpyG/core/graph.py:56-119`` constructor, ``.n/.m/.order``, ``.data``, ``.mapping``,
``edge_to_index``, CSR / CSC views, ``to(device)``).

Differences that matter for speed, not for results:
* ``edge_to_index`` (a Python dict over all edges, 8.4 us per edge in the reference,
  ``core/graph.py:110-112``) and the CSR / CSC views are built on first access;
* layers produced by ``aggregate_edge_index`` are already (row, col)-sorted and validated on the
  device, so ``Graph._from_sorted`` skips the sort and the validation pass.
``degrees`` / ``transition_probabilities`` (graph.py:486-533) feed the model-selection statistics of
``MultiOrderModel`` and run on the device (``csrc/selection.cu``); the other generic graph utilities of the
reference (Laplacian, ``__add__`` ...) are out of scope.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _staging, ops
from .data import Data, EdgeIndex
from .index_map import IndexMap


class Graph:
    def __init__(self, data: Data, mapping: IndexMap | None = None):
        self.mapping = mapping if mapping is not None else IndexMap()
        if "num_nodes" not in data and "edge_index" in data:
            ei = data.edge_index
            data.num_nodes = int(ei.max()) + 1 if ei.numel() else 0
        if not isinstance(data.edge_index, EdgeIndex):
            data.edge_index = EdgeIndex(data.edge_index, sparse_size=(data.num_nodes, data.num_nodes))
        if (data.edge_index.get_sparse_size(0) != data.num_nodes
                or data.edge_index.get_sparse_size(1) != data.num_nodes):
            raise ValueError("sparse size of EdgeIndex must match number of nodes!")
        self.data = data
        # stable sort by row, carried over to the edge attributes (graph.py:103-105)
        data.edge_index, perm = data.edge_index.sort_by("row")
        if perm is not None:
            for attr in self.edge_attrs():
                if attr != "edge_index":
                    data[attr] = data[attr][perm]
        data.edge_index.validate()
        self._finish()

    @classmethod
    def _from_sorted(cls, data: Data, mapping: IndexMap | None = None) -> "Graph":
        """For layers whose edge_index is known to be (row, col)-sorted and in range."""
        g = cls.__new__(cls)
        g.mapping = mapping if mapping is not None else IndexMap()
        if not isinstance(data.edge_index, EdgeIndex):
            data.edge_index = EdgeIndex(data.edge_index, sparse_size=(data.num_nodes, data.num_nodes), sort_order="row")
        g.data = data
        g._finish()
        return g

    def _finish(self) -> None:
        self._edge_to_index = None
        if "node_sequence" not in self.data:
            self.data.node_sequence = torch.arange(self.data.num_nodes, device=self.data.edge_index.device).reshape(-1, 1)

    # ---- lazily built views -------------------------------------------------------------
    @property
    def edge_to_index(self) -> dict:
        if self._edge_to_index is None:
            rows, cols = self.data.edge_index.as_tensor().cpu().tolist()
            self._edge_to_index = {(r, c): i for i, (r, c) in enumerate(zip(rows, cols))}
        return self._edge_to_index

    @property
    def row_ptr(self) -> torch.Tensor:
        return self.data.edge_index.get_csr()[0][0]

    @property
    def col(self) -> torch.Tensor:
        return self.data.edge_index.get_csr()[0][1]

    @property
    def col_ptr(self) -> torch.Tensor:
        return self.data.edge_index.get_csc()[0][0]

    @property
    def row(self) -> torch.Tensor:
        return self.data.edge_index.get_csc()[0][1]

    # ---- constructors --------------------------------------------------------------------
    @staticmethod
    def from_edge_index(edge_index: torch.Tensor, mapping: IndexMap | None = None, num_nodes: int | None = None) -> "Graph":
        if not num_nodes:
            return Graph(Data(edge_index=edge_index), mapping=mapping)
        if mapping is not None and mapping.num_ids() != num_nodes:
            raise ValueError("Number of node IDs in mapping must match num_nodes")
        return Graph(Data(edge_index=edge_index, num_nodes=num_nodes), mapping=mapping)

    @staticmethod
    def from_edge_list(edge_list, is_undirected: bool = False, mapping: IndexMap | None = None, device=None) -> "Graph":
        if len(edge_list) == 0:
            return Graph(Data(edge_index=torch.empty((2, 0), dtype=torch.long, device=device), num_nodes=0), mapping=IndexMap())
        if mapping is None:
            node_ids = np.unique(np.array(edge_list))
            if np.issubdtype(node_ids.dtype, np.str_) and np.char.isnumeric(node_ids).all():
                node_ids = np.sort(node_ids.astype(int)).astype(str)
            mapping = IndexMap(node_ids)
        n = mapping.num_ids()
        ei = EdgeIndex(mapping.to_idxs(edge_list, device=device).T.contiguous(), sparse_size=(n, n))
        return Graph(Data(edge_index=ei, num_nodes=n), mapping=mapping)

    # ---- accessors -----------------------------------------------------------------------
    @property
    def device(self) -> torch.device:
        return self.data.edge_index.device

    def to(self, device) -> "Graph":
        self.data = self.data.to(device)
        return self

    def node_attrs(self):
        return self.data.node_attrs()

    def edge_attrs(self):
        return self.data.edge_attrs()

    @property
    def n(self) -> int:
        return int(self.data.num_nodes)

    @property
    def m(self) -> int:
        return int(self.data.edge_index.size(1))

    @property
    def order(self) -> int:
        ns = self.data.node_sequence
        return 1 if ns is None else int(ns.size(1))

    @property
    def nodes(self) -> list:
        """IDs (or indices without a mapping) of all nodes, tuples for higher-order graphs (graph.py:329-342)."""
        if self.order > 1:
            seq = self.data.node_sequence.cpu().numpy()
            ids = self.mapping.to_ids(np.arange(self.n)) if isinstance(self.mapping.node_ids, np.ndarray) and \
                self.mapping.id_shape != (-1,) else seq
            return [tuple(v.tolist()) for v in ids]
        ids = self.mapping.to_ids(np.arange(self.n))
        return [str(v) if isinstance(v, np.str_) else (v.item() if isinstance(v, np.generic) else v) for v in ids]

    @property
    def edges(self) -> list:
        nodes = self.nodes
        rows, cols = self.data.edge_index.as_tensor().cpu().tolist()
        return [(nodes[r], nodes[c]) for r, c in zip(rows, cols)]

    def get_successors(self, row_idx: int) -> torch.Tensor:
        ptr = self.row_ptr
        return self.col[ptr[row_idx]: ptr[row_idx + 1]]

    def get_predecessors(self, col_idx: int) -> torch.Tensor:
        ptr = self.col_ptr
        return self.row[ptr[col_idx]: ptr[col_idx + 1]]

    # ---- degrees / transition probabilities (graph.py:486-533) -----------------------------
    def _degree_tensor(self, mode: str, edge_attr: str | None) -> torch.Tensor:
        ei = self.data.edge_index
        dev, to_host = _staging.compute_device(ei)
        t = _staging.up(ei, dev).as_subclass(torch.Tensor)
        w = None
        if edge_attr:
            w = getattr(self.data, edge_attr, None)
            if w is None:
                raise AttributeError(f"edge attribute {edge_attr} not found")
            w = _staging.up(w, dev)
        if mode == "in":
            grouped = ops.csc_build(t, self.n, self.n)          # stable: slots keep the edge order inside a target
            ptr, perm = grouped.colptr, grouped.eid
        else:
            ptr, perm = ops.sorted_ids_ptr(t[0], self.n), None  # Graph keeps edge_index sorted by row
        if w is None:
            d = (ptr[1:] - ptr[:-1]).to(torch.int32)            # torch_geometric.utils.degree(..., dtype=torch.int)
        else:
            d = ops.segment_sum(ptr, w, perm)
            if w.dtype.is_floating_point and w.dtype != torch.float32:
                d = d.to(w.dtype)
        return _staging.down(d, to_host)

    def degrees(self, mode: str = "in", edge_attr: str | None = None, return_tensor: bool = False):
        d = self._degree_tensor(mode, edge_attr)
        if return_tensor:
            return d
        return {node: degree.item() for node, degree in zip(self.nodes, d)}

    def in_degrees(self) -> dict:
        return self.degrees(mode="in")

    def out_degrees(self) -> dict:
        return self.degrees(mode="out")

    def transition_probabilities(self, edge_attr: str | None = None) -> torch.Tensor:
        """edge weight / (weighted) out-degree of the edge's source, one value per edge (graph.py:518-533)."""
        ei = self.data.edge_index
        dev, to_host = _staging.compute_device(ei)
        t = _staging.up(ei, dev).as_subclass(torch.Tensor)
        w = _staging.up(getattr(self.data, edge_attr, None), dev) if edge_attr is not None else None
        ptr = ops.sorted_ids_ptr(t[0], self.n)
        denom = ops.segment_sum(ptr, w)
        return _staging.down(ops.edge_ratio(t[0], w, denom), to_host)

    def is_edge(self, v, w) -> bool:
        return (self.mapping.to_idx(v), self.mapping.to_idx(w)) in self.edge_to_index

    def __str__(self) -> str:
        return f"Directed graph with {self.n} nodes and {self.m} edges"
