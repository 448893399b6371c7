"""``Graph`` -- the output container of the lift path, with the reference's surface
(``src/pathpyG/core/graph.py:56-119`` constructor, ``.n/.m/.order``, ``.data``, ``.mapping``,
``edge_to_index``, CSR / CSC views, ``to(device)``).

Differences that matter for speed, not for results:
* ``edge_to_index`` (a Python dict over all edges, 8.4 us per edge in the reference,
  ``core/graph.py:110-112``) and the CSR / CSC views are built on first access;
* layers produced by ``aggregate_edge_index`` are already (row, col)-sorted and validated on the
  device, so ``Graph._from_sorted`` skips the sort and the validation pass.
``degrees`` / ``transition_probabilities`` (graph.py:486-533) feed the model-selection statistics of
``MultiOrderModel`` and run on the device (``csrc/selection.cu``); ``to_undirected`` / ``to_weighted_graph``
(graph.py:211-270) merge their edges with the library's coalesce (radix sort + run reduction).  The remaining
members (``successors``, ``__getitem__``, ``__add__``, ``laplacian`` ...) are index plumbing around those and keep
the reference's semantics, quirks included (``__add__`` concatenates ``node_sequence``, graph.py:742).
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
                val = data[attr]
                data[attr] = val[perm] if isinstance(val, torch.Tensor) else val[perm.cpu().numpy()]
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
        ei = EdgeIndex(mapping.to_idxs(edge_list, device=device).T.contiguous(), sparse_size=(n, n),
                       is_undirected=is_undirected)
        return Graph(Data(edge_index=ei, num_nodes=n), mapping=mapping)

    # ---- conversions (graph.py:211-270) ------------------------------------------------------
    def _coalesce(self, edge_index: torch.Tensor, weight: torch.Tensor | None, reduce: str):
        """(row, col)-sorted distinct edges with reduced weights, on the CUDA path, results on the input's device."""
        dev, to_host = _staging.compute_device(edge_index)
        ei, w = ops.coalesce(_staging.up(edge_index, dev), None, self.n, _staging.up(weight, dev), reduce)
        return _staging.down(ei, to_host), _staging.down(w, to_host)

    def to_undirected(self) -> "Graph":
        """Every edge in both directions, duplicates merged; an edge attribute of a merged edge is the one of the
        lowest-numbered original edge (PyG ``to_undirected(..., reduce="min")`` over the edge numbers, graph.py:225-251)."""
        ei = self.data.edge_index.as_tensor()
        both = torch.cat([ei, ei.flip(0)], dim=1)
        number = torch.arange(ei.size(1), device=ei.device).repeat(2)
        new_ei, attr_idx = self._coalesce(both, number, "min")
        data = Data(edge_index=EdgeIndex(new_ei, sparse_size=(self.n, self.n), is_undirected=True), num_nodes=self.n)
        for attr in self.node_attrs():
            data[attr] = self.data[attr]
        for attr in self.edge_attrs():
            val = self.data[attr]
            data[attr] = val[attr_idx] if isinstance(val, torch.Tensor) else val[attr_idx.cpu().numpy()]
        return Graph(data, self.mapping)

    def to_weighted_graph(self) -> "Graph":
        """Multi-edges merged into one edge whose ``edge_weight`` counts them (graph.py:253-270)."""
        ei, w = self._coalesce(self.data.edge_index.as_tensor(), None, "sum")
        return Graph(Data(edge_index=ei, edge_weight=w, num_nodes=self.n), mapping=self.mapping)

    # ---- accessors -----------------------------------------------------------------------
    @property
    def device(self) -> torch.device:
        return self.data.edge_index.device

    def to(self, device) -> "Graph":
        self.data = self.data.to(device)
        return self

    def node_attrs(self) -> list:
        """Names of the node-level attributes: ``node_*`` except ``node_sequence`` (graph.py:294-309)."""
        return [k for k in self.data.keys() if k != "node_sequence" and k.startswith("node_")]

    def edge_attrs(self) -> list:
        """Names of the edge-level attributes: ``edge_*`` except ``edge_index`` (graph.py:311-327)."""
        return [k for k in self.data.keys() if k != "edge_index" and k.startswith("edge_")]

    @property
    def n(self) -> int:
        return int(self.data.num_nodes)

    @property
    def m(self) -> int:
        """Number of edges; an undirected graph counts (a,b),(b,a) once (graph.py:646-662)."""
        if self.is_directed():
            return int(self.data.edge_index.size(1))
        ei = self.data.edge_index.as_tensor()
        loops = int((ei[0] == ei[1]).sum())
        return int((ei.size(1) - loops) / 2 + loops)

    def is_directed(self) -> bool:
        return not self.data.edge_index.is_undirected

    def is_undirected(self) -> bool:
        return self.data.edge_index.is_undirected

    def has_self_loops(self) -> bool:
        return self.data.has_self_loops()

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
        if row_idx + 1 < ptr.size(0):
            return self.col[ptr[row_idx]: ptr[row_idx + 1]]
        return torch.tensor([], device=self.device)

    def get_predecessors(self, col_idx: int) -> torch.Tensor:
        ptr = self.col_ptr
        if col_idx + 1 < ptr.size(0):
            return self.row[ptr[col_idx]: ptr[col_idx + 1]]
        return torch.tensor([], device=self.device)

    def _ids_of(self, idxs: torch.Tensor) -> list:
        ids = self.mapping.to_ids(idxs)
        out = ids.tolist()
        return list(map(tuple, out)) if self.order > 1 else out

    def successors(self, node) -> list:
        """IDs (indices without a mapping) of the nodes ``node`` points to (graph.py:393-412)."""
        return self._ids_of(self.get_successors(self.mapping.to_idx(node)))

    def predecessors(self, node) -> list:
        return self._ids_of(self.get_predecessors(self.mapping.to_idx(node)))

    def is_edge(self, v, w) -> bool:
        """Membership of ``w`` in the CSR row of ``v`` (graph.py:433-449)."""
        row = self.mapping.to_idx(v)
        ptr = self.row_ptr
        return bool((self.col[ptr[row]: ptr[row + 1]] == self.mapping.to_idx(w)).any())

    # ---- matrices (graph.py:451-470, 535-566): host-side scipy objects -----------------------------
    def sparse_adj_matrix(self, edge_attr=None):
        """``scipy.sparse.coo_matrix`` in edge order (PyG ``to_scipy_sparse_matrix``)."""
        import scipy.sparse

        ei = self.data.edge_index.as_tensor().cpu().numpy()
        if edge_attr is None:
            val = np.ones(ei.shape[1])
        else:
            val = self.data[edge_attr]
            val = (val.detach().cpu().numpy() if isinstance(val, torch.Tensor) else np.asarray(val)).reshape(-1)
        return scipy.sparse.coo_matrix((val, (ei[0], ei[1])), (self.n, self.n))

    def laplacian(self, normalization=None, edge_attr: str | None = None):
        """PyG ``get_laplacian`` (self-loops dropped; ``None``: D - A, ``"sym"``: I - D^-1/2 A D^-1/2,
        ``"rw"``: I - D^-1 A; entries = edges, then the diagonal) as a ``scipy.sparse.coo_matrix``."""
        import scipy.sparse

        if normalization not in (None, "sym", "rw"):
            raise ValueError(f"Invalid normalization {normalization}")
        ei = self.data.edge_index.as_tensor()
        w = torch.ones(ei.size(1), device=ei.device) if edge_attr is None else self.data[edge_attr].reshape(-1)
        keep = ei[0] != ei[1]
        ei, w = ei[:, keep], w[keep]
        n = int(ei.max()) + 1 if ei.numel() else 0  # the reference passes no num_nodes (graph.py:553,558)
        deg = torch.zeros(n, dtype=w.dtype, device=w.device).index_add_(0, ei[0], w)
        loops = torch.arange(n, device=ei.device)
        if normalization is None:
            val = torch.cat([-w, deg])
        elif normalization == "sym":
            dis = deg.pow(-0.5)
            dis.masked_fill_(dis == float("inf"), 0)
            val = torch.cat([-(dis[ei[0]] * w * dis[ei[1]]), torch.ones(n, dtype=w.dtype, device=w.device)])
        else:
            dinv = 1.0 / deg
            dinv.masked_fill_(dinv == float("inf"), 0)
            val = torch.cat([-(dinv[ei[0]] * w), torch.ones(n, dtype=w.dtype, device=w.device)])
        idx = torch.cat([ei, loops.repeat(2, 1)], dim=1).cpu().numpy()
        return scipy.sparse.coo_matrix((val.cpu().numpy(), (idx[0], idx[1])), (n, n))

    # ---- attribute access (graph.py:568-633) ---------------------------------------------------------
    def __getitem__(self, key):
        if not isinstance(key, tuple):
            if key in self.data.keys():
                return self.data[key]
            raise KeyError(key + " is not a graph attribute")
        if key[0] in self.node_attrs():
            return self.data[key[0]][self.mapping.to_idx(key[1])]
        if key[0] in self.edge_attrs():
            return self.data[key[0]][self.edge_to_index[self.mapping.to_idx(key[1]), self.mapping.to_idx(key[2])]]
        raise KeyError(key[0] + " is not a node or edge attribute")

    def __setitem__(self, key, val) -> None:
        if not isinstance(key, tuple):
            if key.startswith("node_") and val.size(0) != self.n:
                raise ValueError("Attribute must have same length as number of nodes")
            if key.startswith("edge_") and val.size(0) != self.m:
                raise ValueError("Attribute must have same length as number of edges")
            self.data[key] = val
        elif key[0].startswith("node_") or key[0].startswith("edge_"):
            if key[0] not in self.data.keys():
                raise KeyError("Attribute does not yet exist. Setting the value of a specific node attribute"
                               + "requires that the attribute already exists.")
            if key[0].startswith("node_"):
                self.data[key[0]][self.mapping.to_idx(key[1])] = val
            else:
                self.data[key[0]][self.edge_to_index[self.mapping.to_idx(key[1]), self.mapping.to_idx(key[2])]] = val
        else:
            raise KeyError("node and edge specific attributes should be prefixed with 'node_' or 'edge_'")

    # ---- union of two graphs (graph.py:688-770) ---------------------------------------------------------
    def __add__(self, other: "Graph", reduce: str = "sum") -> "Graph":
        d1, m1 = self.data.clone(), self.mapping
        d2, m2 = other.data.clone(), other.mapping
        ids1, ids2 = m1.to_ids(np.arange(self.n)), m2.to_ids(np.arange(other.n))
        mapping = IndexMap(np.unique(np.concatenate([ids1, ids2]), axis=0).tolist())
        dev = d1.edge_index.device
        d1.edge_index = mapping.to_idxs(m1.to_ids(d1.edge_index.as_tensor()), device=dev)
        d2.edge_index = mapping.to_idxs(m2.to_ids(d2.edge_index.as_tensor()), device=d2.edge_index.device)
        d = d1.concat(d2)
        d.num_nodes = mapping.num_ids()
        d.edge_index = EdgeIndex(d.edge_index, sparse_size=(d.num_nodes, d.num_nodes))
        if "inverse_idx" in d:  # higher-order layers: edge -> node numbers follow the new numbering
            d.inverse_idx = mapping.to_idxs(np.concatenate([m1.to_ids(d1.inverse_idx), m2.to_ids(d2.inverse_idx)]),
                                            device=d.inverse_idx.device)
        where = mapping.to_idxs(np.concatenate([ids1, ids2]))
        for k in d1.keys():
            if k != "node_sequence" and k.startswith("node_"):
                if not isinstance(d[k], torch.Tensor):
                    raise ValueError("Node attribute " + k + " is not a tensor and cannot be reduced.")
                d[k] = _scatter_rows(d[k], where.to(d[k].device), d.num_nodes, reduce)
        return Graph(d, mapping=mapping)

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
        elif w.dtype == torch.float32:
            d = ops.segment_sum(ptr, w, perm)
        else:
            # the reference's scatter(..., reduce="sum") keeps the weight dtype (graph.py:506,512): integer weights are
            # summed exactly in int64, float64 weights in float64 (differences of one running sum over the grouped order)
            grouped_w = w if perm is None else w[perm.long()]
            acc = torch.float64 if w.dtype.is_floating_point else torch.int64
            run = torch.cat([torch.zeros(1, dtype=acc, device=dev), torch.cumsum(grouped_w.to(acc), 0)])
            d = (run[ptr[1:].long()] - run[ptr[:-1].long()]).to(w.dtype)
        return _staging.down(d, to_host)

    def degrees(self, mode: str = "in", edge_attr: str | None = None, return_tensor: bool = False):
        d = self._degree_tensor(mode, edge_attr)
        if return_tensor:
            return d
        return {node: degree.item() for node, degree in zip(self.nodes, d)}

    @property
    def in_degrees(self) -> dict:
        return self.degrees(mode="in")

    @property
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

    def __str__(self) -> str:
        """Head line + attribute summary in the reference's format (graph.py:772-805)."""
        from pprint import pformat

        def kind(v):
            return str(torch.Tensor) + " -> " + str(v.size()) if isinstance(v, torch.Tensor) else str(type(v))

        def per_item(v, count):  # PyG Data.is_node_attr / is_edge_attr: leading dimension matches
            return isinstance(v, (torch.Tensor, np.ndarray)) and v.ndim >= 1 and v.shape[0] == count

        head = "Undirected" if self.is_undirected() else "Directed"
        info: dict = {"Node Attributes": {}, "Edge Attributes": {}, "Graph Attributes": {}}
        node_attrs, edge_attrs = self.node_attrs(), self.edge_attrs()
        num_edges = int(self.data.edge_index.size(1))
        for k in self.data.keys():
            v = self.data[k]
            if k in node_attrs:
                info["Node Attributes"][k] = kind(v)
            elif k in edge_attrs:
                info["Edge Attributes"][k] = kind(v)
            elif k != "edge_index" and not (per_item(v, self.n) and "edge" not in k) \
                    and not (per_item(v, num_edges) and "node" not in k):
                info["Graph Attributes"][k] = kind(v)
        return f"{head} graph with {self.n} nodes and {self.m} edges\n" + pformat(info, indent=4, width=160)


def _scatter_rows(src: torch.Tensor, index: torch.Tensor, dim_size: int, reduce: str) -> torch.Tensor:
    """``torch_geometric.utils.scatter(src, index, dim=0, dim_size, reduce)`` for the node attributes of ``__add__``."""
    shape = (dim_size, *src.shape[1:])
    if reduce in ("sum", "add"):
        return torch.zeros(shape, dtype=src.dtype, device=src.device).index_add_(0, index, src)
    name = {"mean": "mean", "mul": "prod", "min": "amin", "max": "amax"}.get(reduce)
    if name is None:
        raise ValueError(f"Encountered invalid `reduce` argument '{reduce}'")
    idx = index.reshape(-1, *([1] * (src.dim() - 1))).expand_as(src)
    init = torch.ones(shape, dtype=src.dtype, device=src.device) if reduce == "mul" else \
        torch.zeros(shape, dtype=src.dtype, device=src.device)
    return init.scatter_reduce_(0, idx, src, name, include_self=reduce == "mul")
