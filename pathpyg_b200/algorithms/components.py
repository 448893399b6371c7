"""Connected components of a static graph (reference ``src/pathpyG/algorithms/components.py``): scipy on the
host-side adjacency matrix of ``Graph.sparse_adj_matrix`` -- utilities around the containers, not part of the GPU path."""
from __future__ import annotations

import numpy as np
from scipy.sparse.csgraph import connected_components as _scipy_components

from ..core.graph import Graph


def connected_components(graph: Graph, connection: str = "weak"):
    """``(number of components, label of every node)``; ``connection`` is ``"weak"`` or ``"strong"``."""
    count, labels = _scipy_components(graph.sparse_adj_matrix(), directed=graph.is_directed(), connection=connection,
                                      return_labels=True)
    return count, labels


def largest_connected_component(graph: Graph, connection: str = "weak") -> Graph:
    """The sub-graph spanned by the component with the most nodes (the first such label on ties), rebuilt from its
    edge list like the reference does -- so its mapping holds exactly the nodes that keep an edge."""
    _, labels = connected_components(graph, connection)
    sizes = np.bincount(labels)
    candidates = np.flatnonzero(sizes == sizes.max())
    first_seen = [int(np.argmax(labels == c)) for c in candidates]   # ties: the component met first in node order
    biggest = int(candidates[int(np.argmin(first_seen))])
    ei = graph.data.edge_index.as_tensor().cpu().numpy()
    keep = (labels[ei[0]] == biggest) & (labels[ei[1]] == biggest)
    ids = graph.mapping.to_ids(ei[:, keep])
    edges = list(zip(ids[0].tolist(), ids[1].tolist()))
    return Graph.from_edge_list(edges, is_undirected=graph.is_undirected())
