"""One-dimensional Weisfeiler-Leman colour refinement as an isomorphism heuristic (reference
``src/pathpyG/algorithms/weisfeiler_leman.py``): a caller of ``Graph.__add__`` and ``Graph.successors``; host utility."""
from __future__ import annotations

from ..core.graph import Graph


def WeisfeilerLeman_test(g1: Graph, g2: Graph, features_g1: dict | None = None, features_g2: dict | None = None):
    """``(may be isomorphic, colours of g1's nodes, colours of g2's nodes)``.  ``False`` proves the graphs are not
    isomorphic; ``True`` means the refinement found no difference.  Both graphs need node ids, and the ids must not
    overlap.  Colours are numbered in the order they are first met while sweeping the nodes of ``g1 + g2``."""
    if g1.mapping is None or g2.mapping is None:
        raise Exception("Graphs must contain IndexMap that assigns node IDs")
    if set(g1.mapping.node_ids).intersection(g2.mapping.node_ids):
        raise Exception("node identifiers of graphs must not overlap")
    union = g1 + g2
    nodes = union.nodes
    neighbours = {v: union.successors(v) for v in nodes}
    if features_g1 is None or features_g2 is None:
        colour = {v: "0" for v in nodes}
    else:
        colour = {**features_g1, **features_g2}
    palette: dict = {}
    while True:
        refined = {}
        for v in nodes:
            signature = (colour[v], tuple(sorted(colour[w] for w in neighbours[v])))
            refined[v] = palette.setdefault(signature, len(palette) + 1)
        if len(set(colour.values())) == len(set(refined.values())):   # no class was split: stable
            break
        colour = refined
    c1, c2 = [colour[v] for v in g1.nodes], [colour[v] for v in g2.nodes]
    return sorted(c1) == sorted(c2), c1, c2
