"""Run the reference's OWN source for the lift path in this container.

TEST INFRASTRUCTURE ONLY.  Used by ``tests/golden/make_golden.py`` (to generate the
committed fixtures) and by ``tests/test_oracle_vs_reference.py`` (skipped when
``/root/reference`` is absent, i.e. on the GPU box).

``import pathpyG`` fails here because torch_geometric is not installed
(``core/graph.py:18-22``).  The arithmetic of the lift path, however, sits in two of the
reference's own files; this loader executes them unmodified from where they lie with
stand-ins for their imports:

* ``src/pathpyG/algorithms/lift_order.py``  (all four functions)
* ``src/pathpyG/algorithms/temporal.py``    (``lift_order_temporal``)

Stand-ins: ``torch_geometric.utils.{degree,cumsum,coalesce}`` -> ``oracle/pyg.py``;
``torch_geometric.data.Data`` -> attribute bag; ``pathpyG.core.graph.Graph`` -> a holder that
applies the stable row sort of ``Graph.__init__`` (graph.py:103-105) and nothing else.
No reference source is copied into this repository.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types
from contextlib import contextmanager

import torch

from . import pyg

REFERENCE_ROOT = os.environ.get("PATHPYG_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src/pathpyG/algorithms/lift_order.py"))


class _Data:
    def __init__(self, **kw):
        self.__dict__.update(kw)

    def __contains__(self, key):
        return key in self.__dict__

    def __getitem__(self, key):
        return self.__dict__[key]


class _Graph:
    """Holder standing in for pathpyG.core.graph.Graph: keeps ``data`` and applies the
    stable sort by row that the real constructor applies to edge_index and edge attrs."""

    def __init__(self, data, mapping=None):
        order = torch.sort(data.edge_index[0], stable=True).indices
        data.edge_index = data.edge_index[:, order]
        if getattr(data, "edge_weight", None) is not None:
            data.edge_weight = data.edge_weight[order]
        self.data = data
        self.mapping = mapping

    @staticmethod
    def from_edge_index(edge_index, mapping=None, num_nodes=None):
        return _Graph(_Data(edge_index=edge_index, num_nodes=num_nodes), mapping)


def _coalesce(edge_index, edge_attr=None, num_nodes=None, reduce="sum"):
    return pyg.coalesce(edge_index, edge_attr, num_nodes, reduce)


@contextmanager
def _stand_ins():
    names = {}

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        names[name] = m
        return m

    mod("torch_geometric")
    mod("torch_geometric.data", Data=_Data)
    mod("torch_geometric.utils", degree=lambda index, num_nodes=None, dtype=None: pyg.degree(index, num_nodes, dtype or torch.float),
        cumsum=pyg.cumsum, coalesce=_coalesce)
    mod("pathpyG", Graph=_Graph)
    mod("pathpyG.core")
    mod("pathpyG.core.graph", Graph=_Graph)
    mod("pathpyG.core.temporal_graph", TemporalGraph=object)
    mod("pathpyG.utils", to_numpy=lambda t: t.numpy())
    saved = {k: sys.modules.get(k) for k in names}
    sys.modules.update(names)
    try:
        yield
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def _exec_file(rel_path: str, private_name: str):
    path = os.path.join(REFERENCE_ROOT, rel_path)
    spec = importlib.util.spec_from_file_location(private_name, path)
    module = importlib.util.module_from_spec(spec)
    with _stand_ins():
        spec.loader.exec_module(module)
    return module


_cache: dict = {}


def lift_order_module():
    """The reference's ``algorithms/lift_order.py`` executed as a module."""
    if "lift" not in _cache:
        _cache["lift"] = _exec_file("src/pathpyG/algorithms/lift_order.py", "_ref_lift_order")
    return _cache["lift"]


def temporal_module():
    """The reference's ``algorithms/temporal.py`` executed as a module."""
    if "temporal" not in _cache:
        m = _exec_file("src/pathpyG/algorithms/temporal.py", "_ref_temporal")
        m.tqdm = lambda it, *a, **k: it  # silence the progress bar only
        _cache["temporal"] = m
    return _cache["temporal"]


class TemporalInput:
    """What ``lift_order_temporal`` reads from a TemporalGraph: ``g.data.edge_index`` / ``g.data.time``."""

    def __init__(self, edge_index: torch.Tensor, time: torch.Tensor):
        self.data = _Data(edge_index=edge_index, time=time)


def ref_lift_order_temporal(edge_index, time, delta):
    return temporal_module().lift_order_temporal(TemporalInput(edge_index, time), delta)
