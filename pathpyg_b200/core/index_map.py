"""Node-ID <-> index bookkeeping with the reference's ``IndexMap`` surface
(``src/pathpyG/core/index_map.py``), plus a lazy higher-order variant.

The reference builds the mapping of every higher-order layer eagerly with one Python iteration
and one ``.cpu()`` per higher-order node (``core/multi_order_model.py:119,177-179``; 7 us per
node).  ``HigherOrderIndexMap`` keeps the device-side ``node_sequence`` instead and materialises
IDs / the reverse dictionary only when somebody asks, with one vectorised look-up.
"""
from __future__ import annotations

import numpy as np
import torch


def _as_numpy(x) -> np.ndarray:
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy()
    return np.asarray(x)


class IndexMap:
    def __init__(self, node_ids=None) -> None:
        self._node_ids: np.ndarray | None = None
        self._id_to_idx: dict = {}
        self.id_shape: tuple = (-1,)
        if node_ids is not None:
            self.add_ids(node_ids)

    # ---- storage (properties so that the lazy subclass can defer them)
    @property
    def node_ids(self):
        return self._node_ids

    @node_ids.setter
    def node_ids(self, value):
        self._node_ids = value

    @property
    def id_to_idx(self) -> dict:
        return self._id_to_idx

    @property
    def has_ids(self) -> bool:
        return self.node_ids is not None

    def num_ids(self) -> int:
        return 0 if self.node_ids is None else len(self.node_ids)

    def _key(self, v):
        return tuple(v.tolist()) if self.id_shape != (-1,) else (v.item() if isinstance(v, np.generic) else v)

    def add_id(self, node_id) -> None:
        key = tuple(node_id) if isinstance(node_id, (list, tuple)) else node_id
        if key in self.id_to_idx:
            raise ValueError("ID already present in the mapping.")
        if isinstance(node_id, (list, tuple)):
            arr = _as_numpy(node_id)
            self.id_shape = (-1, *arr.shape)
            arr = arr.reshape(1, *arr.shape)
        else:
            arr = _as_numpy([node_id])
        idx = self.num_ids()
        self._node_ids = arr if self._node_ids is None else np.concatenate((self._node_ids, arr))
        self._id_to_idx[key] = idx

    def add_ids(self, node_ids) -> None:
        start = self.num_ids()
        if isinstance(node_ids, list) and len(node_ids) and isinstance(node_ids[0], (list, tuple)):
            self.id_shape = (-1, *_as_numpy(node_ids[0]).shape)
        new = _as_numpy(node_ids)
        if new.size == 0:
            return
        merged = new if self._node_ids is None else np.concatenate((self._node_ids, new))
        distinct = np.unique(merged, axis=0 if self.id_shape != (-1,) else None)
        if len(distinct) != len(merged):
            raise ValueError("IDs are not unique or already present in the mapping.")
        self._node_ids = merged
        if self.id_shape != (-1,):
            self._id_to_idx.update({tuple(v): start + i for i, v in enumerate(new.tolist())})
        else:
            self._id_to_idx.update({v: start + i for i, v in enumerate(new.tolist())})

    def _sorted_order(self) -> np.ndarray:
        cached = getattr(self, "_order_cache", None)
        if cached is None or cached[0] != len(self.node_ids):
            cached = (len(self.node_ids), np.argsort(self.node_ids, kind="stable"))
            self._order_cache = cached
        return cached[1]

    def to_id(self, idx: int):
        if not self.has_ids:
            return idx
        if self.id_shape == (-1,):
            v = self.node_ids[idx]
            return str(v) if self.node_ids.dtype.type is np.str_ else v
        return tuple(self.node_ids[idx].tolist())

    def to_ids(self, idxs):
        if self.node_ids is None:
            return idxs
        return self.node_ids[_as_numpy(idxs)]

    def to_idx(self, node):
        if not self.has_ids:
            return node
        return self.id_to_idx[tuple(node) if self.id_shape != (-1,) else node]

    def to_idxs(self, nodes, device=None) -> torch.Tensor:
        if not self.has_ids:
            return torch.as_tensor(_as_numpy(nodes) if not isinstance(nodes, torch.Tensor) else nodes, device=device)
        arr = _as_numpy(nodes)
        lut = self.id_to_idx
        if self.id_shape == (-1,):
            if arr.size >= 64 and arr.dtype.kind == self.node_ids.dtype.kind and arr.dtype.kind in "iufUS":
                # bulk look-up: one binary search per id over the sorted ids instead of one dictionary probe
                # (reference core/index_map.py:368); unknown ids raise KeyError like the dictionary would
                ids = self.node_ids
                order = self._sorted_order()
                pos = np.searchsorted(ids, arr.reshape(-1), sorter=order)
                pos[pos == len(ids)] = 0
                hit = order[pos]
                bad = ids[hit] != arr.reshape(-1)
                if bad.any():
                    raise KeyError(arr.reshape(-1)[bad][0].item())
                return torch.from_numpy(hit.astype(np.int64)).to(device).reshape(arr.shape)
            flat = [lut[v] for v in arr.reshape(-1).tolist()]
            return torch.tensor(flat, device=device).reshape(arr.shape)
        k = len(self.id_shape) - 1
        lead = arr.shape[: arr.ndim - k]
        flat = [lut[tuple(v) if isinstance(v, list) else v] for v in arr.reshape(-1, *arr.shape[arr.ndim - k:]).tolist()]
        return torch.tensor(flat, device=device).reshape(lead)

    def __eq__(self, other) -> bool:
        if not isinstance(other, IndexMap):
            return NotImplemented
        a, b = self.node_ids, other.node_ids
        if a is None or b is None:
            return a is None and b is None
        return a.shape == b.shape and bool((a == b).all())

    __hash__ = None

    def __str__(self) -> str:
        if not self.has_ids:
            return ""
        return "\n".join(f"{self.to_id(i)} -> {i}" for i in range(self.num_ids())) + "\n"


class HigherOrderIndexMap(IndexMap):
    """IDs of higher-order nodes = tuples of first-order IDs, derived on demand from the layer's
    ``node_sequence`` ([n, k] indices into ``base``).  Compares equal to the eager
    ``IndexMap([tuple(base.to_ids(v)) for v in node_sequence])`` of the reference."""

    def __init__(self, base: IndexMap, node_sequence: torch.Tensor) -> None:
        super().__init__()
        self._base = base
        self._node_sequence = node_sequence
        self.id_shape = (-1, int(node_sequence.size(1)))
        self._materialised = False
        self._dict_ready = False

    @property
    def node_ids(self):
        if not self._materialised:
            seq = self._node_sequence.detach().cpu().numpy()
            self._node_ids = self._base.to_ids(seq) if self._base.has_ids else seq
            if isinstance(self._node_ids, torch.Tensor):
                self._node_ids = self._node_ids.numpy()
            self._materialised = True
        return self._node_ids

    @node_ids.setter
    def node_ids(self, value):
        self._node_ids = value
        self._materialised = True

    @property
    def id_to_idx(self) -> dict:
        if not self._dict_ready:
            self._id_to_idx = {tuple(v): i for i, v in enumerate(self.node_ids.tolist())}
            self._dict_ready = True
        return self._id_to_idx

    def num_ids(self) -> int:
        return int(self._node_sequence.size(0)) if not self._materialised else len(self._node_ids)

    @property
    def has_ids(self) -> bool:
        return True
