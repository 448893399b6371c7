"""Minimal stand-ins for ``torch_geometric.data.Data`` and ``torch_geometric.EdgeIndex``.

torch_geometric is a dependency of the reference but not of this package: the hot path only
needs an attribute container and an edge-index tensor that remembers its sparse size and
caches its CSR / CSC views.  Only the members the reference's hot path touches are provided
(``core/graph.py:79-119``, ``core/temporal_graph.py:50-75``, ``core/multi_order_model.py``).
"""
from __future__ import annotations

import copy

import torch


class EdgeIndex(torch.Tensor):
    """``[2, E]`` int64 tensor with a sparse size; every torch op on it yields a plain tensor."""

    __torch_function__ = torch._C._disabled_torch_function_impl

    @staticmethod
    def __new__(cls, data, sparse_size=None, sort_order=None, is_undirected=False, **kwargs):
        t = torch.as_tensor(data, **kwargs)
        if isinstance(t, EdgeIndex):
            t = t.as_subclass(torch.Tensor)
        if t.dtype != torch.int64:
            t = t.long()
        out = torch.Tensor._make_subclass(cls, t)
        out._sparse_size = tuple(sparse_size) if sparse_size is not None else (None, None)
        out._sort_order = sort_order
        out._is_undirected = bool(is_undirected)
        out._cache = {}
        return out

    # ---- PyG-compatible surface
    def as_tensor(self) -> torch.Tensor:
        return self.as_subclass(torch.Tensor)

    @property
    def sparse_size(self):
        return self._sparse_size

    def get_sparse_size(self, dim: int | None = None):
        """Like PyG: a dimension that was not declared is the largest index of that row + 1."""
        if dim is None:
            return (self.get_sparse_size(0), self.get_sparse_size(1))
        size = self._sparse_size[dim]
        if size is None:
            row = self.as_tensor()[dim]
            size = int(row.max()) + 1 if row.numel() else 0
        return size

    @property
    def sort_order(self):
        return self._sort_order

    @property
    def is_undirected(self) -> bool:
        """The flag the creator set (PyG keeps it as metadata; it is not derived from the edges)."""
        return self._is_undirected

    def validate(self) -> "EdgeIndex":
        t = self.as_tensor()
        if t.dim() != 2 or t.size(0) != 2:
            raise ValueError(f"'EdgeIndex' needs to have a shape of [2, *] (got {list(t.shape)})")
        if t.numel() > 0:
            lo, hi = int(t.min()), int(t.max())
            if lo < 0:
                raise ValueError(f"'EdgeIndex' contains negative indices (got {lo})")
            n0, n1 = self._sparse_size
            if n0 is not None and int(t[0].max()) >= n0:
                raise ValueError(f"'EdgeIndex' contains larger indices than its number of rows (got {int(t[0].max())}, but expected values smaller than {n0})")
            if n1 is not None and int(t[1].max()) >= n1:
                raise ValueError(f"'EdgeIndex' contains larger indices than its number of columns (got {int(t[1].max())}, but expected values smaller than {n1})")
        return self

    def sort_by(self, sort_order: str = "row", stable: bool = True):
        """Returns ``(sorted EdgeIndex, permutation)``; the permutation is None if already sorted."""
        if self._sort_order == sort_order:
            return self, None
        t = self.as_tensor()
        key = t[0] if sort_order == "row" else t[1]
        if key.numel() < 2 or bool((key[1:] >= key[:-1]).all()):
            out = EdgeIndex(t, sparse_size=self._sparse_size, sort_order=sort_order, is_undirected=self._is_undirected)
            return out, None
        perm = torch.sort(key, stable=True).indices
        return EdgeIndex(t[:, perm], sparse_size=self._sparse_size, sort_order=sort_order,
                         is_undirected=self._is_undirected), perm

    def _ptr(self, ids: torch.Tensor, n: int) -> torch.Tensor:
        counts = torch.bincount(ids, minlength=n)
        ptr = counts.new_zeros(n + 1)
        torch.cumsum(counts, 0, out=ptr[1:])
        return ptr

    def get_csr(self):
        """((rowptr, col), perm) -- requires / establishes row order."""
        if "csr" not in self._cache:
            t = self.as_tensor()
            perm = None
            if self._sort_order != "row":
                perm = torch.sort(t[0], stable=True).indices
                t = t[:, perm]
            n = self._sparse_size[0] if self._sparse_size[0] is not None else (int(t[0].max()) + 1 if t.numel() else 0)
            self._cache["csr"] = ((self._ptr(t[0], n), t[1]), perm)
        return self._cache["csr"]

    def get_csc(self):
        """((colptr, row), perm)."""
        if "csc" not in self._cache:
            t = self.as_tensor()
            perm = None
            if self._sort_order != "col":
                perm = torch.sort(t[1], stable=True).indices
                t = t[:, perm]
            n = self._sparse_size[1] if self._sparse_size[1] is not None else (int(t[1].max()) + 1 if t.numel() else 0)
            self._cache["csc"] = ((self._ptr(t[1], n), t[0]), perm)
        return self._cache["csc"]

    def to(self, *args, **kwargs):  # keep the wrapper across device moves
        return EdgeIndex(self.as_tensor().to(*args, **kwargs), sparse_size=self._sparse_size, sort_order=self._sort_order,
                         is_undirected=self._is_undirected)

    def __repr__(self):
        return f"EdgeIndex({self.as_tensor().tolist()}, sparse_size={self._sparse_size})"

    def __deepcopy__(self, memo):
        return EdgeIndex(self.as_tensor().clone(), sparse_size=self._sparse_size, sort_order=self._sort_order,
                         is_undirected=self._is_undirected)


class Data:
    """Attribute bag with the few ``torch_geometric.data.Data`` methods the path uses."""

    def __init__(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)

    # ---- mapping protocol
    def __contains__(self, key) -> bool:
        return key in self.__dict__ and self.__dict__[key] is not None

    def __getitem__(self, key):
        return self.__dict__[key]

    def __setitem__(self, key, value):
        self.__dict__[key] = value

    def __getattr__(self, key):
        # like PyG: unknown attributes read as None
        if key.startswith("__"):
            raise AttributeError(key)
        return None

    def keys(self):
        return [k for k, v in self.__dict__.items() if v is not None]

    def to_dict(self):
        return {k: self.__dict__[k] for k in self.keys()}

    # ---- sizes
    @property
    def num_edges(self) -> int:
        ei = self.__dict__.get("edge_index")
        return 0 if ei is None else int(ei.size(1))

    def has_self_loops(self) -> bool:
        ei = self.__dict__.get("edge_index")
        return bool(ei is not None and ei.numel() and (ei[0] == ei[1]).any())

    def concat(self, other: "Data") -> "Data":
        """``torch_geometric.data.Data.concat``: tensors of both objects joined along their cat dimension
        (last for ``*index*`` keys, first otherwise), numpy arrays along axis 0; other values are kept from ``self``."""
        import numpy as np

        out = copy.copy(self)
        for k, v in self.__dict__.items():
            w = other.__dict__.get(k)
            if isinstance(v, torch.Tensor) and isinstance(w, torch.Tensor):
                a, b = v.as_subclass(torch.Tensor), w.as_subclass(torch.Tensor)
                out.__dict__[k] = torch.cat([a, b], dim=-1 if "index" in k else 0)
            elif isinstance(v, np.ndarray) and isinstance(w, np.ndarray):
                out.__dict__[k] = np.concatenate([v, w])
        return out

    def _per_item(self, k, v, count) -> bool:
        """PyG's shape test (``BaseStorage.is_node_attr`` / ``is_edge_attr``): a tensor or numpy array with at least
        one dimension whose concatenation dimension (last for ``*index*`` keys, first otherwise) has ``count`` entries."""
        import numpy as np

        if not isinstance(v, (torch.Tensor, np.ndarray)) or v.ndim == 0 or count is None:
            return False
        return v.shape[-1 if "index" in k else 0] == count

    def edge_attrs(self):
        """Keys holding one entry per edge, by PyG's rule (``GlobalStorage.is_edge_attr``, used by the reference through
        ``data.edge_attrs()``): the size along the concatenation dimension equals ``num_edges``; when ``num_nodes ==
        num_edges`` makes that ambiguous, the name decides ('edge' in the key; ``time`` is per event by definition).
        Numpy arrays count too (the reference keeps string attributes as numpy arrays, io/pandas.py:91-92)."""
        m, n = self.num_edges, self.__dict__.get("num_nodes")
        out = []
        for k, v in self.__dict__.items():
            if k == "edge_index":
                out.append(k)
            elif self._per_item(k, v, m) and (n != m or "edge" in k or k == "time"):
                out.append(k)
        return out

    def node_attrs(self):
        """Keys holding one entry per node (``GlobalStorage.is_node_attr``): size ``num_nodes`` along the concatenation
        dimension; with ``num_nodes == num_edges`` keys naming an edge are excluded."""
        m, n = self.num_edges, self.__dict__.get("num_nodes")
        return [k for k, v in self.__dict__.items()
                if k != "edge_index" and self._per_item(k, v, n) and (n != m or ("edge" not in k and k != "time"))]

    # ---- time order (TemporalGraph input, multi_order_model.py:148-151)
    def is_sorted_by_time(self) -> bool:
        t = self.__dict__.get("time")
        return t is None or t.numel() < 2 or bool((t[1:] >= t[:-1]).all())

    def sort_by_time(self) -> "Data":
        perm = torch.sort(self.time, stable=True).indices
        out = copy.copy(self)
        for k in self.edge_attrs():
            v = self.__dict__[k]
            if k == "edge_index":
                out.__dict__[k] = v[:, perm]
            else:
                out.__dict__[k] = v[perm] if isinstance(v, torch.Tensor) else v[perm.cpu().numpy()]
        return out

    # ---- device
    def to(self, device, non_blocking: bool = False) -> "Data":
        out = copy.copy(self)
        for k, v in self.__dict__.items():
            if isinstance(v, EdgeIndex):
                out.__dict__[k] = v.to(device)
            elif isinstance(v, torch.Tensor):
                out.__dict__[k] = v.to(device, non_blocking=non_blocking)
        return out

    def clone(self) -> "Data":
        return copy.deepcopy(self)

    def __repr__(self):
        parts = []
        for k in self.keys():
            v = self.__dict__[k]
            parts.append(f"{k}={list(v.shape)}" if isinstance(v, torch.Tensor) else f"{k}={v}")
        return "Data(" + ", ".join(parts) + ")"
