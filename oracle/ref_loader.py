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

import numpy as np
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


# ----------------------------------------------------------------------------------------------
# model-selection statistics: the reference's own method bodies (multi_order_model.py:243-509)
# ----------------------------------------------------------------------------------------------
class _RefEdgeIndex(torch.Tensor):
    """Stand-in for torch_geometric.EdgeIndex with the two members get_mon_dof touches
    (``sort_by("row")`` and ``matmul``, multi_order_model.py:298-301); unit edge values."""

    __torch_function__ = torch._C._disabled_torch_function_impl

    @staticmethod
    def wrap(t, num_nodes):
        out = torch.Tensor._make_subclass(_RefEdgeIndex, t)
        out.num_nodes = num_nodes
        return out

    def sort_by(self, order):
        assert order == "row"
        t = self.as_subclass(torch.Tensor)
        perm = torch.sort(t[0], stable=True).indices
        return _RefEdgeIndex.wrap(t[:, perm], self.num_nodes), perm

    def matmul(self, other):
        import warnings

        n = self.num_nodes
        a, b = self.as_subclass(torch.Tensor), other.as_subclass(torch.Tensor)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            A = torch.sparse_coo_tensor(a, torch.ones(a.size(1)), (n, n))
            B = torch.sparse_coo_tensor(b, torch.ones(b.size(1)), (n, n))
            C = torch.sparse.mm(A, B).coalesce()
        return _RefEdgeIndex.wrap(C.indices(), n), C.values()


class _RefLayer:
    """What the statistics read from ``self.layers[k]``: ``.data`` fields and ``transition_probabilities`` -- the
    latter executes the reference's own ``Graph.degrees`` / ``Graph.transition_probabilities`` bodies."""

    def __init__(self, layer, graph_methods):
        self.data = _Data(edge_index=_RefEdgeIndex.wrap(layer.edge_index, layer.num_nodes), num_nodes=layer.num_nodes,
                          num_edges=layer.edge_index.size(1), edge_weight=layer.edge_weight, inverse_idx=layer.inverse_idx,
                          node_sequence=layer.node_sequence)
        self.n = layer.num_nodes
        self._m = graph_methods

    def degrees(self, *a, **k):
        return self._m["degrees"](self, *a, **k)

    def transition_probabilities(self, *a, **k):
        return self._m["transition_probabilities"](self, *a, **k)


def _methods_of(rel_path: str, class_name: str, names, env: dict) -> dict:
    """Compile the named methods of one class from a reference file, unmodified, into ``env``."""
    import ast

    path = os.path.join(REFERENCE_ROOT, rel_path)
    tree = ast.parse(open(path).read(), filename=path)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == class_name)
    picked = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in names]
    module = ast.Module(body=picked, type_ignores=[])
    exec(compile(module, path, "exec"), env)
    return {n: env[n] for n in names}


def selection_methods():
    """``(model_factory, methods)``: ``model_factory(layers)`` builds the object the reference's statistics see as
    ``self``; ``methods`` maps name -> the reference's own function (call as ``methods[name](model, ...)``)."""
    if "selection" not in _cache:
        import logging
        import typing

        from scipy.stats import chi2

        tg = types.SimpleNamespace(utils=types.SimpleNamespace(
            degree=lambda index, num_nodes=None, dtype=None: pyg.degree(index, num_nodes, dtype or torch.float)))
        genv = dict(torch=torch, torch_geometric=tg, scatter=lambda src, index, dim=0, dim_size=None, reduce="sum":
                    pyg.scatter(src, index, dim_size, reduce), Union=typing.Union, Dict=typing.Dict, Any=typing.Any)
        graph_methods = _methods_of("src/pathpyG/core/graph.py", "Graph", ["degrees", "transition_probabilities"], genv)
        menv = dict(torch=torch, cumsum=pyg.cumsum, chi2=chi2, Optional=typing.Optional, Data=_Data, PathData=object,
                    logger=logging.getLogger("ref"), lift_order_edge_index=lift_order_module().lift_order_edge_index)
        names = ["get_mon_dof", "get_zeroth_order_log_likelihood", "get_intermediate_order_log_likelihood",
                 "get_mon_log_likelihood", "likelihood_ratio_test"]
        methods = _methods_of("src/pathpyG/core/multi_order_model.py", "MultiOrderModel", names, menv)

        class Model:
            def __init__(self, layers):
                self.layers = {k: _RefLayer(v, graph_methods) for k, v in layers.items()}

        for name, fn in methods.items():
            setattr(Model, name, fn)
        _cache["selection"] = (Model, methods)
    return _cache["selection"]


def ref_walks_data(walks):
    """The ``dag_graph`` argument of the statistics: the fields of ``PathData.data``."""
    return _Data(edge_index=walks.edge_index, node_sequence=walks.node_sequence, dag_weight=walks.dag_weight,
                 dag_num_edges=walks.dag_num_edges, dag_num_nodes=walks.dag_num_nodes,
                 num_nodes=int(walks.node_sequence.size(0)))


# ----------------------------------------------------------------------------------------------
# ingest: the reference's own io/pandas.py, core/index_map.py and core/path_data.py
# ----------------------------------------------------------------------------------------------
class _RefTemporalGraph:
    """Stand-in for pathpyG.TemporalGraph in io/pandas.py: applies the time ordering of the real constructor
    (temporal_graph.py:58-63) with a STABLE sort (the reference's argsort leaves the order of ties unspecified)."""

    def __init__(self, data, mapping=None):
        order = torch.sort(data.time, stable=True).indices
        m = data.edge_index.size(1)
        for k, v in list(data.__dict__.items()):
            if k == "edge_index":
                data.edge_index = v[:, order]
            elif isinstance(v, torch.Tensor) and v.dim() >= 1 and v.size(0) == m and (k == "time" or k.startswith("edge_")):
                data.__dict__[k] = v[order]
            elif isinstance(v, np.ndarray) and k.startswith("edge_") and v.shape[0] == m:
                data.__dict__[k] = v[order.numpy()]
        self.data, self.mapping = data, mapping
        self.n, self.m = data.num_nodes, m


class _SettableData(_Data):
    def __setitem__(self, key, value):
        self.__dict__[key] = value


def io_module():
    """The reference's ``io/pandas.py`` executed as a module, with its own ``IndexMap`` and ``PathData``."""
    if "io" not in _cache:
        import numpy  # noqa: F401

        to_numpy = lambda t: t.numpy() if isinstance(t, torch.Tensor) else np.asarray(t)  # noqa: E731
        names = {}

        def mod(name, **attrs):
            m = types.ModuleType(name)
            m.__dict__.update(attrs)
            names[name] = m
            return m

        mod("torch_geometric")
        mod("torch_geometric.data", Data=_SettableData)
        mod("torch_geometric.utils", cumsum=pyg.cumsum)
        mod("pathpyG")
        mod("pathpyG.utils", to_numpy=to_numpy)
        mod("pathpyG.utils.convert", to_numpy=to_numpy)
        mod("pathpyG.core")
        mod("pathpyG.core.graph", Graph=_Graph)
        mod("pathpyG.core.temporal_graph", TemporalGraph=_RefTemporalGraph)
        saved = {k: sys.modules.get(k) for k in list(names) + ["pathpyG.core.index_map", "pathpyG.core.path_data"]}
        sys.modules.update(names)
        try:
            for rel, name in (("src/pathpyG/core/index_map.py", "pathpyG.core.index_map"),
                              ("src/pathpyG/core/path_data.py", "pathpyG.core.path_data"),
                              ("src/pathpyG/io/pandas.py", "_ref_io_pandas")):
                spec = importlib.util.spec_from_file_location(name, os.path.join(REFERENCE_ROOT, rel))
                module = importlib.util.module_from_spec(spec)
                sys.modules[name] = module
                spec.loader.exec_module(module)
            _cache["io"] = sys.modules.pop("_ref_io_pandas")
        finally:
            for k, v in saved.items():
                if v is None:
                    sys.modules.pop(k, None)
                else:
                    sys.modules[k] = v
    return _cache["io"]


# ----------------------------------------------------------------------------------------------
# containers: the reference's own edge-merging members of Graph / TemporalGraph
# ----------------------------------------------------------------------------------------------
class _KwEdgeIndex(torch.Tensor):
    """Stand-in for torch_geometric.EdgeIndex as the container members use it: ``EdgeIndex(data=, sparse_size=,
    is_undirected=)``, ``as_tensor()``, ``size``, ``max``, column indexing, ``device``."""

    __torch_function__ = torch._C._disabled_torch_function_impl

    @staticmethod
    def __new__(cls, data, sparse_size=None, is_undirected=False):
        out = torch.Tensor._make_subclass(cls, torch.as_tensor(data).as_subclass(torch.Tensor))
        out.sparse_size_, out.is_undirected = sparse_size, is_undirected
        return out

    def as_tensor(self):
        return self.as_subclass(torch.Tensor)

    def __getitem__(self, key):  # PyG keeps the wrapper when columns are selected
        out = self.as_subclass(torch.Tensor)[key]
        return _KwEdgeIndex(out, self.sparse_size_, self.is_undirected) if out.dim() == 2 else out


class _ContainerSelf:
    """``self`` of the extracted methods: ``data``, ``mapping``, ``m``, ``node_attrs()``, ``edge_attrs()``."""

    def __init__(self, data):
        self.data, self.mapping = data, None

    @property
    def m(self):
        return self.data.edge_index.size(1)

    def node_attrs(self):  # graph.py:294-309
        return [k for k in self.data.__dict__ if k != "node_sequence" and k.startswith("node_")]

    def edge_attrs(self):  # graph.py:311-327
        return [k for k in self.data.__dict__ if k != "edge_index" and k.startswith("edge_")]


def container_methods():
    """``(make_self, methods)``: the reference's ``Graph.to_undirected`` / ``Graph.to_weighted_graph``
    (core/graph.py:211-270) and ``TemporalGraph.to_static_graph`` (core/temporal_graph.py:191-220), compiled
    unmodified; PyG's ``to_undirected`` / ``coalesce`` are the restatements of ``oracle/pyg.py`` and ``Graph(...)`` is
    the holder that applies the constructor's stable row sort."""
    if "containers" not in _cache:
        import typing

        class Holder(_Graph):
            def __init__(self, data, mapping=None):
                order = torch.sort(data.edge_index.as_subclass(torch.Tensor)[0], stable=True).indices
                for k, v in list(data.__dict__.items()):
                    if k == "edge_index":
                        data.edge_index = v.as_subclass(torch.Tensor)[:, order]
                    elif k.startswith("edge_") and isinstance(v, torch.Tensor):
                        data.__dict__[k] = v[order]
                self.data, self.mapping = data, mapping

            @staticmethod
            def from_edge_index(edge_index, mapping=None, num_nodes=None):
                return Holder(_SettableData(edge_index=edge_index, num_nodes=num_nodes), mapping)

        tg = types.SimpleNamespace(utils=types.SimpleNamespace(
            coalesce=lambda ei, attr=None, num_nodes=None, reduce="sum": pyg.coalesce(
                ei.as_subclass(torch.Tensor), attr, num_nodes if num_nodes is not None else int(ei.max()) + 1, reduce)))
        env = dict(torch=torch, torch_geometric=tg, Data=_SettableData, EdgeIndex=_KwEdgeIndex, Graph=Holder,
                   Optional=typing.Optional, Tuple=typing.Tuple)
        genv = dict(env)
        methods = _methods_of("src/pathpyG/core/graph.py", "Graph", ["to_undirected", "to_weighted_graph"], genv)
        # the method is compiled under the name of the PyG function it calls: rebind the global to the function
        genv["to_undirected"] = lambda ei, edge_attr=None, num_nodes=None, reduce="add": pyg.to_undirected(
            ei.as_subclass(torch.Tensor), edge_attr, num_nodes, reduce)
        methods.update(_methods_of("src/pathpyG/core/temporal_graph.py", "TemporalGraph", ["to_static_graph"], dict(env)))

        def make_self(edge_index, num_nodes=None, **attrs):
            data = _SettableData(edge_index=_KwEdgeIndex(edge_index, sparse_size=(num_nodes, num_nodes)), **attrs)
            data.num_nodes = num_nodes
            data.num_edges = edge_index.size(1)
            return _ContainerSelf(data)

        _cache["containers"] = (make_self, methods)
    return _cache["containers"]


# ----------------------------------------------------------------------------------------------
# drop-in check: reference modules that only need the containers, executed ON this package's containers
# ----------------------------------------------------------------------------------------------
def reference_module_on(package, rel_path: str, name: str):
    """Execute one of the reference's own modules (``rel_path`` under ``src/pathpyG/``) with ``pathpyG.*`` imports
    resolved to ``package`` (= ``pathpyg_b200``): the module's functions then run on this package's ``Graph`` /
    ``TemporalGraph`` / ``PathData``.  Third-party imports the module makes but the compared functions do not need
    (``torch_geometric``, ``tqdm``) are stubbed."""
    stubs = {}

    def mod(modname, **attrs):
        m = types.ModuleType(modname)
        m.__dict__.update(attrs)
        stubs[modname] = m
        return m

    mod("pathpyG", Graph=package.Graph, TemporalGraph=package.TemporalGraph, PathData=package.PathData, IndexMap=package.IndexMap)
    mod("pathpyG.core")
    mod("pathpyG.core.graph", Graph=package.Graph)
    mod("pathpyG.core.temporal_graph", TemporalGraph=package.TemporalGraph)
    mod("pathpyG.core.path_data", PathData=package.PathData)
    mod("pathpyG.core.index_map", IndexMap=package.IndexMap)
    mod("pathpyG.utils", to_numpy=package.utils.to_numpy)
    mod("pathpyG.utils.convert", to_numpy=package.utils.to_numpy)
    mod("torch_geometric.data", Data=package.Data)
    mod("pathpyG.algorithms")
    mod("pathpyG.algorithms.temporal", lift_order_temporal=package.algorithms.lift_order_temporal,
        temporal_shortest_paths=package.algorithms.temporal_shortest_paths)
    mod("torch_geometric")
    mod("torch_geometric.utils", to_networkx=None, degree=pyg.degree)
    mod("tqdm", tqdm=lambda it, *a, **k: it)
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    try:
        spec = importlib.util.spec_from_file_location(name, os.path.join(REFERENCE_ROOT, "src/pathpyG", rel_path))
        module = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(module)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return module
