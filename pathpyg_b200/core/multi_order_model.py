"""``MultiOrderModel`` builders (reference ``src/pathpyG/core/multi_order_model.py``):
``iterate_lift_order`` (:83-122), ``from_temporal_graph`` (:124-192), ``from_path_data`` (:194-241),
``to_dbgnn_data`` (:511-554), and the model-selection statistics that consume the built layers
(SURVEY.md 8f rank 3): ``get_mon_dof`` (:243-312), the log-likelihoods (:314-409),
``likelihood_ratio_test`` (:411-459) and ``estimate_order`` (:461-509).

All tensors of one build stay on the GPU from the first kernel to the last; host inputs are staged
once at entry and the finished layers are moved back once at exit.
"""
from __future__ import annotations

import logging

import torch

from .. import _staging, ops
from ..utils.dbgnn import generate_bipartite_edge_index
from .data import Data, EdgeIndex
from .graph import Graph
from .index_map import HigherOrderIndexMap, IndexMap
from .path_data import PathData
from .temporal_graph import TemporalGraph

logger = logging.getLogger("root")


_copy_streams: dict = {}


def _copy_stream(dev: torch.device) -> torch.cuda.Stream:
    """One side stream per device for uploads that overlap compute."""
    if dev not in _copy_streams:
        _copy_streams[dev] = torch.cuda.Stream(dev)
    return _copy_streams[dev]


def _plain(t):
    return t.as_subclass(torch.Tensor) if isinstance(t, torch.Tensor) else t


def _aggregate(edge_index, node_sequence, edge_weight, aggr="sum") -> Graph:
    """Device-side aggregate_edge_index (lift_order.py:109-152)."""
    unique_nodes, inverse_idx = ops.unique_rows(node_sequence)
    n = int(unique_nodes.size(0))
    remap = node_sequence.reshape(-1) if node_sequence.size(1) == 1 else inverse_idx
    agg_index, agg_weight = ops.coalesce(edge_index, remap, n, edge_weight, aggr)
    data = Data(edge_index=EdgeIndex(agg_index, sparse_size=(n, n), sort_order="row"), num_nodes=n,
                node_sequence=unique_nodes, edge_weight=agg_weight, inverse_idx=inverse_idx)
    return Graph._from_sorted(data)


def _extend_node_sequence(node_sequence, edge_index):
    """k-gram of every line-graph edge: the (k-1)-gram of its source + the last node of its target
    (multi_order_model.py:114,165)."""
    return torch.cat([node_sequence[edge_index[0]], node_sequence[edge_index[1]][:, -1:]], dim=1)


def _lift_step(edge_index, node_sequence, edge_weight, aggr, save):
    n_prev = int(node_sequence.size(0))
    ho_index = ops.lift_order_edge_index(edge_index, n_prev)
    if edge_weight is not None:
        edge_weight = ops.pair_attributes(ho_index, edge_weight, aggr, index_bound=int(edge_index.size(1)))
    node_sequence = _extend_node_sequence(node_sequence, edge_index)
    gk = _aggregate(ho_index, node_sequence, edge_weight) if save else None
    return ho_index, node_sequence, edge_weight, gk


class _LayerChain:
    """Builds the De Bruijn layers of consecutive orders with ONE radix sort per order.

    The reference finds the nodes of layer k with ``torch.unique(node_sequence, dim=0)`` over one k-gram
    row per line-graph edge of level k-1 (lift_order.py:133) and then coalesces the mapped edges
    (:139-144): two sorts per order, the first over [E, k] int64 rows.  But the k-gram of such an edge is
    (the (k-1)-gram of its source) ++ (last node of its target), so two edges carry the same k-gram iff they
    map to the same layer-(k-1) edge, and k-grams ordered lexicographically are ordered like
    (rank of prefix, rank of suffix) = the (row, col) order of the coalesced layer-(k-1) edges.  Hence

        nodes of layer k        = distinct edges of layer k-1, in their (row, col) order,
        inverse_idx of layer k  = the output edge every level-(k-1) edge was merged into,
        node_sequence of layer k= extend_rows(node_sequence of layer k-1, edge_index of layer k-1),

    and the [E, k] row matrix of the reference (multi_order_model.py:114) is never materialised.
    """

    def __init__(self, model: "MultiOrderModel", cached: bool, max_order: int):
        self.model, self.cached, self.max_order = model, cached, max_order
        self.order = 0
        self.edge_index = self.node_sequence = self.inverse_next = None
        self.overlapped = None

    def _store(self, agg_index, agg_weight, n, node_sequence, inverse_idx) -> None:
        if self.cached or self.order == self.max_order:
            data = Data(edge_index=EdgeIndex(agg_index, sparse_size=(n, n), sort_order="row"), num_nodes=n,
                        node_sequence=node_sequence, edge_weight=agg_weight, inverse_idx=inverse_idx)
            self.model.layers[self.order] = Graph._from_sorted(data)

    def first_layer(self, edge_index, remap, n, node_sequence, inverse_idx, edge_weight, overlap=None) -> None:
        """``overlap``: a pending count -> fill operation (the temporal / line-graph lift of the next level), or a
        callable that enqueues one; it is finished right after this layer's sort, under the same synchronisation."""
        self.order = 1
        more = self.max_order > 1
        pending = ops.coalesce_begin(edge_index, remap, n, edge_weight, "sum", return_inverse=more)
        if callable(overlap):
            overlap = overlap()  # enqueued BEHIND the layer sort: whatever it still waits for (a staged copy) is hidden
        res = pending.finish()
        self.overlapped = overlap.finish() if overlap is not None else None
        self._store(res[0], res[1], n, node_sequence, inverse_idx)
        if more:
            # layer-1 ids are the node values themselves, so the 2-grams are the coalesced edges transposed
            self.edge_index, self.inverse_next = res[0], res[2]
            self.node_sequence = None

    def next_layer(self, line_index, edge_weight, overlap=None) -> None:
        """``line_index`` [2, E_k]: the line graph whose nodes are the level-(k-1) edges."""
        self.order += 1
        more = self.order < self.max_order
        if self.node_sequence is None:
            node_sequence = self.edge_index.t().contiguous()
        else:
            node_sequence = ops.extend_rows(self.node_sequence, self.edge_index)
        n = int(node_sequence.size(0))
        inverse_idx = self.inverse_next
        pending = ops.coalesce_begin(line_index, inverse_idx, n, edge_weight, "sum", return_inverse=more)
        res = pending.finish()
        self.overlapped = overlap.finish() if overlap is not None else None
        self._store(res[0], res[1], n, node_sequence, inverse_idx)
        self.edge_index, self.node_sequence = res[0], node_sequence
        self.inverse_next = res[2] if more else None


def _sort_free_ok(edge_index, edge_weight, n, event_graph, same_data, dev) -> bool:
    """The generation-order chain covers the standard call: a device build from the graph's own time-sorted events with
    unit or float32 weights.  Everything else (a caller-supplied event graph, shuffled time stamps, other weight dtypes)
    takes the sort-per-order chain.  ``PPG_CHAIN=0`` switches it off (A/B measurements)."""
    import os
    if os.environ.get("PPG_CHAIN", "1") == "0" or event_graph is not None or not same_data or dev.type != "cuda":
        return False
    if edge_index.dtype != torch.int64 or not 0 < edge_index.size(1) < (1 << 31) or not 0 < n < (1 << 31):
        return False
    return edge_weight is None or (edge_weight.dtype == torch.float32 and edge_weight.dim() == 1)


def _sort_free_layers(model, edge_index, time, delta, n, max_order, edge_weight, cached, grouped_ws, before_time) -> None:
    from ..chain import TemporalChain

    dev = edge_index.device
    ids = torch.arange(n, device=dev)
    state = {"ns": ids.unsqueeze(1), "ei": None}

    def store(order, agg_index, agg_weight, num_nodes, inverse):
        if order == 1:
            node_sequence, inverse = state["ns"], ids
        elif order == 2:
            node_sequence = state["ei"].t().contiguous()
        else:
            node_sequence = ops.extend_rows(state["ns"], state["ei"])
        state["ns"], state["ei"] = node_sequence, agg_index
        if cached or order == max_order:
            data = Data(edge_index=EdgeIndex(agg_index, sparse_size=(num_nodes, num_nodes), sort_order="row"), num_nodes=num_nodes,
                        node_sequence=node_sequence, edge_weight=agg_weight, inverse_idx=inverse)
            model.layers[order] = Graph._from_sorted(data)

    if edge_weight is not None:
        edge_weight = edge_weight.contiguous()
    TemporalChain(edge_index.contiguous(), n, edge_weight, max_order).run(time, delta, store, grouped_ws, before_time, cached)


class MultiOrderModel:
    """Higher-order De Bruijn graph layers keyed by order (``layers: dict[int, Graph]``)."""

    def __init__(self) -> None:
        self.layers: dict[int, Graph] = {}

    def __str__(self) -> str:
        max_order = max(list(self.layers.keys())) if self.layers else 0
        return f"MultiOrderModel with max. order {max_order}"

    def to(self, device) -> "MultiOrderModel":
        for g in self.layers.values():
            g.to(device)
        return self

    # ------------------------------------------------------------------------------------------
    @staticmethod
    def iterate_lift_order(edge_index, node_sequence, mapping: IndexMap, edge_weight=None, aggr: str = "src",
                           save: bool = True):
        """multi_order_model.py:83-122: lift by one order, extend the node sequences, aggregate."""
        if aggr not in ("src", "dst", "max", "mul", "add"):
            raise ValueError(f"Unknown aggregation method {aggr}")
        dev, to_host = _staging.compute_device(edge_index, node_sequence, edge_weight)
        ei, ns, w = (_plain(_staging.up(t, dev)) for t in (edge_index, node_sequence, edge_weight))
        ho_index, ns, w, gk = _lift_step(ei.long(), ns.long(), w, aggr, save)
        if gk is not None:
            if to_host:
                gk.to("cpu")
            gk.mapping = HigherOrderIndexMap(mapping, gk.data.node_sequence)
        return _staging.down(ho_index, to_host), _staging.down(ns, to_host), _staging.down(w, to_host), gk

    # ------------------------------------------------------------------------------------------
    @staticmethod
    def from_temporal_graph(g: TemporalGraph, delta: float | int = 1, max_order: int = 1, weight: str = "edge_weight",
                            cached: bool = True, event_graph: torch.Tensor | None = None, device=None) -> "MultiOrderModel":
        """multi_order_model.py:124-192, one radix sort per order (see ``_LayerChain``).

        ``device`` (extension): build on this CUDA device from a HOST graph and leave the layers there.  The edge
        index is uploaded first; the time stamps follow on a copy stream while the first-order layer is already
        being sorted, and the temporal lift -- the first consumer of the time stamps -- is enqueued behind that sort."""
        m = MultiOrderModel()
        known_sorted = getattr(g, "time_is_known_sorted", lambda: False)()
        data = g.data if known_sorted or g.data.is_sorted_by_time() else g.data.sort_by_time()
        dev, to_host = _staging.compute_device(data.edge_index, data.time)
        staged = device is not None and to_host
        if staged:
            dev, to_host = torch.device(device), False
            if dev.index is None:
                dev = torch.device("cuda", torch.cuda.current_device())
        n = int(data.num_nodes)
        time_ready = rows_ready = grouped_ws = None
        with torch.cuda.device(dev):
            host_ei = _plain(data.edge_index)
            if staged and max_order > 1 and event_graph is None and data is g.data and host_ei.dtype == torch.int64 \
                    and host_ei.size(1) > 0 and n > 0:
                # upload pipeline: source row (main stream) -> group the events by source while the target row and then
                # the time stamps arrive on the copy stream; every consumer waits for exactly the rows it reads
                main, side = torch.cuda.current_stream(dev), _copy_stream(dev)
                edge_index = torch.empty(host_ei.shape, dtype=torch.int64, device=dev)
                edge_index[0].copy_(host_ei[0], non_blocking=True)
                side.wait_stream(main)                   # the copies follow each other: each keeps the full link rate
                edge_index.record_stream(side)
                with torch.cuda.stream(side):
                    edge_index[1].copy_(host_ei[1], non_blocking=True)
                    rows_ready = torch.cuda.Event()
                    rows_ready.record(side)
                    time_dev = _staging.up(data.time, dev)
                    time_ready = torch.cuda.Event()
                    time_ready.record(side)
                grouped_ws = ops.lift_order_temporal_group(edge_index, n)
                main.wait_event(rows_ready)
            else:
                edge_index = _plain(_staging.up(data.edge_index, dev)).long()
        edge_weight = _staging.up(data[weight], dev) if weight in data else None  # None == ones(m), :154-157

        if _sort_free_ok(edge_index, edge_weight, n, event_graph, data is g.data, dev):
            # one device, time-sorted stream: the layers are generated in their final order (csrc/chain.cu)
            _sort_free_layers(m, edge_index, _staging.up(data.time, dev) if time_ready is None else time_dev, delta, n, max_order,
                              edge_weight, cached, grouped_ws,
                              (lambda: (main.wait_event(time_ready), time_dev.record_stream(main))) if time_ready is not None else None)
            m._finish(g.mapping, to_host)
            return m

        chain = _LayerChain(m, cached, max_order)
        pending = None
        if max_order > 1 and event_graph is None:
            # the reference passes `g`, not the locally re-sorted data (:167); identical unless the
            # caller shuffled time stamps after construction, in which case `g.data` is what counts
            src_ei = edge_index if data is g.data else _plain(_staging.up(g.data.edge_index, dev)).long()
            if time_ready is not None:
                def pending():  # called by the chain once the layer-1 sort is enqueued
                    main.wait_event(time_ready)
                    time_dev.record_stream(main)
                    return ops.lift_order_temporal_begin(src_ei, time_dev, delta, n, grouped_ws=grouped_ws)
            else:
                src_t = _staging.up(data.time if data is g.data else g.data.time, dev)
                # count pass runs with the layer-1 sort; a shuffled g.data takes the order-independent route
                pending = ops.lift_order_temporal_begin(src_ei, src_t, delta, n, assume_sorted=data is g.data)
        ids = torch.arange(n, device=dev)
        chain.first_layer(edge_index, None, n, ids.unsqueeze(1), ids, edge_weight, overlap=pending)
        if max_order > 1:
            line_index = chain.overlapped if event_graph is None else _plain(_staging.up(event_graph, dev)).long()
            if edge_weight is not None:
                edge_weight = ops.pair_attributes(line_index, edge_weight, "src",
                                                  index_bound=int(edge_index.size(1)) if event_graph is None else None)
            num_line_nodes = edge_index.size(1)
            for k in range(2, max_order + 1):
                pending = ops.lift_order_edge_index_begin(line_index, num_line_nodes) if k < max_order else None
                chain.next_layer(line_index, edge_weight, overlap=pending)
                if k < max_order:
                    nxt = chain.overlapped
                    if edge_weight is not None:
                        edge_weight = ops.pair_attributes(nxt, edge_weight, "src", index_bound=int(line_index.size(1)))
                    num_line_nodes, line_index = line_index.size(1), nxt
        m._finish(g.mapping, to_host)
        return m

    # ------------------------------------------------------------------------------------------
    @staticmethod
    def from_path_data(path_data: PathData, max_order: int = 1, mode: str = "propagation",
                       cached: bool = True) -> "MultiOrderModel":
        """multi_order_model.py:194-241, one radix sort per order (see ``_LayerChain``)."""
        m = MultiOrderModel()
        pg = path_data.data
        dev, to_host = _staging.compute_device(pg.edge_index, pg.node_sequence)
        edge_index = _plain(_staging.up(pg.edge_index, dev)).long()
        node_sequence = _plain(_staging.up(pg.node_sequence, dev)).long()
        edge_weight = ops.repeat_by_count(_staging.up(pg.dag_weight, dev), _staging.up(pg.dag_num_edges, dev))   # :217
        if mode == "diffusion":
            outdeg = ops.bincount(edge_index[0], node_sequence.size(0))
            edge_weight = edge_weight / outdeg[edge_index[0]]
            aggr = "mul"
        elif mode == "propagation":
            aggr = "src"
        else:
            raise ValueError(f"Unknown mode {mode}")  # the reference dies with a NameError here (:218-224)

        chain = _LayerChain(m, cached, max_order)
        unique_nodes, inverse_idx = ops.unique_rows(node_sequence)
        # first order: the VALUES of node_sequence are the node ids (lift_order.py:135-136)
        chain.first_layer(edge_index, node_sequence.reshape(-1), int(unique_nodes.size(0)), unique_nodes, inverse_idx,
                          edge_weight)
        line_index, num_line_nodes = edge_index, node_sequence.size(0)
        for _ in range(2, max_order + 1):
            nxt = ops.lift_order_edge_index(line_index, num_line_nodes)
            edge_weight = ops.pair_attributes(nxt, edge_weight, aggr, index_bound=int(line_index.size(1)))
            num_line_nodes, line_index = line_index.size(1), nxt
            chain.next_layer(line_index, edge_weight)
        m._finish(path_data.mapping, to_host)
        return m

    def _finish(self, mapping, to_host: bool) -> None:
        for k, layer in self.layers.items():
            if to_host:
                layer.to("cpu")
            layer.mapping = mapping if k == 1 else HigherOrderIndexMap(mapping, layer.data.node_sequence)

    # ------------------------------------------------------------------------------------------
    def get_mon_dof(self, max_order: int | None = None, assumption: str = "paths") -> int:
        """multi_order_model.py:243-312.  Under "paths" the reference materialises every line graph up to
        ``max_order`` only to read its number of columns (:285-291) and multiplies sparse adjacency powers only to
        count their non-empty rows (:294-303).  Both are counts of walks: #columns of the k-th lift = number of
        walks with k edges, non-empty rows of A^k = nodes that start one.  One pass per order over the first-order
        edges computes c_k[v] = sum over out-edges (v, w) of c_{k-1}[w] in exact integers (``ops.walk_counts``)."""
        if max_order is None:
            max_order = max(self.layers)
        if max_order > max(self.layers):
            logger.error("max_order cannot be larger than maximum order of multi-order network")
            raise ValueError("max_order cannot be larger than maximum order of multi-order network")
        n = int(self.layers[1].data.num_nodes)
        dof = n - 1
        if assumption == "paths":
            if max_order >= 1:
                ei = self.layers[1].data.edge_index
                dev, _ = _staging.compute_device(ei)
                walks, sources = ops.walk_counts(_plain(_staging.up(ei, dev)), n, max_order)
                dof += sum(walks) - sum(sources)
        elif assumption == "ngrams":
            for order in range(1, max_order + 1):
                dof += (n ** order) * (n - 1)
        else:
            logger.error("Unknown assumption %s. Only 'path' and 'ngram' are accepted.", assumption)
            raise ValueError(f"Unknown assumption {assumption}. Only 'path' and 'ngram' are accepted.")
        return int(dof)

    @staticmethod
    def _node_counts(node_sequence: torch.Tensor) -> torch.Tensor:
        """``torch.unique(node_sequence, return_counts=True)[1]`` (:335): occurrence counts of the node ids that
        occur, in ascending id order -- ids that never occur are SKIPPED, as in the reference (:336-337)."""
        ids = node_sequence.reshape(-1)
        counts = ops.bincount(ids)
        return counts if bool((counts > 0).all()) else counts[counts > 0]

    def get_zeroth_order_log_likelihood(self, dag_graph: Data) -> float:
        """multi_order_model.py:314-339: sum over walks of weight * log(relative frequency of the start node)."""
        dev, _ = _staging.compute_device(dag_graph.edge_index, dag_graph.node_sequence)
        ei = _plain(_staging.up(dag_graph.edge_index, dev))
        ns = _plain(_staging.up(dag_graph.node_sequence, dev)).reshape(-1)
        is_start = torch.ones(int(dag_graph.num_nodes), dtype=torch.bool, device=dev)
        is_start[ei[1]] = False                                                   # :329-330
        start_nodes = ns[is_start]                                                # :331
        counts = self._node_counts(ns)                                            # :335
        prob = counts / counts.sum()                                              # :338 (int64 / int64 -> float32)
        return ops.weighted_log_sum(_staging.up(dag_graph.dag_weight, dev), prob, start_nodes)    # :339

    def get_intermediate_order_log_likelihood(self, dag_graph: Data, order: int) -> float:
        """multi_order_model.py:341-369: the first transition of every walk in layer ``order``, found through
        ``inverse_idx`` of layer ``order + 1`` at the walk's first position in the order-``order`` line graph."""
        dev, _ = _staging.compute_device(dag_graph.dag_weight, self.layers[order].data.edge_index)
        freq = _staging.up(dag_graph.dag_weight, dev)
        shrunk = _staging.up(dag_graph.dag_num_nodes, dev) - order                # :354-356
        keep = shrunk > 0
        lengths = shrunk[keep]                                                    # :358
        freq = freq[keep]                                                         # :359
        starts = ops.counts_to_offsets(lengths)[0][:-1]                           # cumsum(.)[:-1] of the zero-prefixed sum, :361
        prob = self.layers[order].transition_probabilities()                      # unweighted, as in the reference (:363)
        inverse = _staging.up(self.layers[order + 1].data.inverse_idx, dev)
        return ops.weighted_log_sum(freq, _staging.up(prob, dev), starts, inverse)    # :363-369

    def get_mon_log_likelihood(self, dag_graph: Data, max_order: int = 1) -> float:
        """multi_order_model.py:371-409."""
        if max_order > 0:
            llh = self.get_zeroth_order_log_likelihood(dag_graph)                 # :386
            for order in range(1, max_order):
                llh += self.get_intermediate_order_log_likelihood(dag_graph, order)   # :389-390
            layer = self.layers[max_order]
            dev, _ = _staging.compute_device(layer.data.edge_index)
            prob = layer.transition_probabilities(edge_attr="edge_weight")        # :394
            llh += ops.weighted_log_sum(_staging.up(layer.data.edge_weight, dev), _staging.up(prob, dev))   # :395-397
            return llh
        # zeroth-order model: weighted node frequencies (:402-407)
        dev, _ = _staging.compute_device(dag_graph.node_sequence, dag_graph.dag_weight)
        ns = _plain(_staging.up(dag_graph.node_sequence, dev)).reshape(-1)
        w = ops.repeat_by_count(_staging.up(dag_graph.dag_weight, dev), _staging.up(dag_graph.dag_num_nodes, dev))
        n_ids = int(ns.max()) + 1 if ns.numel() else 0
        # torch.bincount(ids, weights) adds the weights of a node in position order: group the positions by node
        # (stable) and sum every group in slot order
        grouped = ops.csc_build(torch.stack([torch.arange(ns.numel(), device=dev), ns]), ns.numel(), n_ids)
        counts = ops.segment_sum(grouped.colptr, w, grouped.eid)                  # :403-405
        prob = counts / counts.sum()                                              # :406
        return ops.weighted_log_sum(counts, prob)                                 # :407

    def likelihood_ratio_test(self, dag_graph: Data, max_order_null: int = 0, max_order: int = 1,
                              assumption: str = "paths", significance_threshold: float = 0.01) -> tuple:
        """multi_order_model.py:411-459."""
        from scipy.stats import chi2

        if max_order_null >= max_order:
            logger.error("order of null hypothesis must be smaller than order of alternative hypothesis")
            raise ValueError("order of null hypothesis must be smaller than order of alternative hypothesis")
        if max_order > max(self.layers):
            logger.error("order of hypotheses must be smaller than max. order of MultiOrderModel")
            raise ValueError(f"order of hypotheses ({max_order_null} and {max_order}) must be smaller than max. order of "
                             f"MultiOrderModel {max(self.layers)}")
        x = -2 * (self.get_mon_log_likelihood(dag_graph, max_order=max_order_null)
                  - self.get_mon_log_likelihood(dag_graph, max_order=max_order))
        dof_diff = self.get_mon_dof(max_order, assumption=assumption) - self.get_mon_dof(max_order_null, assumption=assumption)
        p = 1 - chi2.cdf(x, dof_diff)
        return (p < significance_threshold), p

    def estimate_order(self, dag_data: PathData, max_order: int | None = None, significance_threshold: float = 0.01) -> int:
        """multi_order_model.py:461-509."""
        if max_order is None:
            max_order = max(self.layers)
        if max_order > max(self.layers):
            logger.error("max_order cannot be larger than maximum order of multi-order network")
            raise ValueError("max_order cannot be larger than maximum order of multi-order network")
        if max_order <= 1:
            logger.error("max_order must be larger than one")
            raise ValueError("max_order must be larger than one")
        if not set(dag_data.mapping.node_ids).issubset(set(self.layers[1].mapping.node_ids)):
            logger.error("Input paths do not have same set of nodes as multi-order network")
            raise ValueError("Input paths do not have same set of nodes as multi-order network")
        max_accepted_order = 1
        dag_graph = dag_data.data
        for k in range(2, max_order + 1):
            if self.likelihood_ratio_test(dag_graph, max_order_null=k - 1, max_order=k,
                                          significance_threshold=significance_threshold)[0]:
                max_accepted_order = k
        return max_accepted_order

    # ------------------------------------------------------------------------------------------
    def to_dbgnn_data(self, max_order: int = 2, mapping: str = "last", x_h: torch.Tensor | None = None) -> Data:
        """multi_order_model.py:511-554.  ``x`` is ``layers[1].data.x`` or a dense one-hot matrix;
        ``x_h`` is the reference's dense ``eye(n_ho)`` unless given explicitly (an extension: the
        one-hot matrix needs n_ho^2 floats and cannot be built for large layers)."""
        if max_order not in self.layers:
            logger.error("Higher-order graph of specified order not found.")
            raise ValueError(f"Higher-order graph of order {max_order} not found.")
        g, gk = self.layers[1], self.layers[max_order]
        dev = g.data.edge_index.device
        x = g.data.x if g.data.x is not None else torch.eye(g.data.num_nodes, g.data.num_nodes, device=dev)
        if x_h is None:
            x_h = torch.eye(gk.data.num_nodes, gk.data.num_nodes, device=gk.data.edge_index.device)
        return Data(
            num_nodes=g.data.num_nodes,
            num_ho_nodes=gk.data.num_nodes,
            x=x,
            x_h=x_h,
            edge_index=g.data.edge_index,
            edge_index_higher_order=gk.data.edge_index,
            edge_weights=g.data.edge_weight.float(),
            edge_weights_higher_order=gk.data.edge_weight.float(),
            bipartite_edge_index=generate_bipartite_edge_index(g, gk, mapping=mapping, device=dev),
            y=g.data.y,
        )
