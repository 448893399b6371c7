"""``PathData`` -- store of observed walks as one concatenated DAG
(``src/pathpyG/core/path_data.py:45-204``): ``data.edge_index`` over global positions,
``node_sequence [sum(L), 1]``, ``dag_weight``, ``dag_num_edges``, ``dag_num_nodes``.

``append_index_walks`` is the bulk entry for walks that are already index tensors on their
device (5M walks in BASELINE config 4): no per-walk Python, no dictionary look-ups.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import ops
from .data import Data
from .index_map import IndexMap


class PathData:
    def __init__(self, mapping: IndexMap | None = None, device=None) -> None:
        self.mapping = mapping if mapping else IndexMap()
        self.data = Data(
            edge_index=torch.empty((2, 0), dtype=torch.long, device=device),
            node_sequence=torch.empty((0, 1), dtype=torch.long, device=device),
            dag_weight=torch.empty(0, dtype=torch.float, device=device),
            dag_num_edges=torch.empty(0, dtype=torch.long, device=device),
            dag_num_nodes=torch.empty(0, dtype=torch.long, device=device),
        )
        self.data.num_nodes = 0

    @property
    def num_paths(self) -> int:
        return len(self.data.dag_num_edges)

    def _append_data(self, edge_index, node_sequence, weights, num_edges, num_nodes) -> None:
        d = self.data
        d.edge_index = torch.cat([d.edge_index, edge_index + d.num_nodes], dim=1)
        d.node_sequence = torch.cat([d.node_sequence, node_sequence])
        d.dag_weight = torch.cat([d.dag_weight, weights.to(d.dag_weight.dtype)])
        d.dag_num_edges = torch.cat([d.dag_num_edges, num_edges])
        d.dag_num_nodes = torch.cat([d.dag_num_nodes, num_nodes])
        d.num_nodes += int(num_nodes.sum())

    def to(self, device) -> "PathData":
        self.data = self.data.to(device)
        return self

    def append_walk(self, node_seq, weight: float = 1.0) -> None:
        dev = self.data.edge_index.device
        idx_seq = self.mapping.to_idxs(node_seq, device=dev).unsqueeze(1)
        pos = torch.arange(len(node_seq), device=dev)
        chain = torch.stack([pos[:-1], pos[1:]])
        self._append_data(chain, idx_seq, torch.tensor([weight], device=dev),
                          torch.tensor([chain.shape[1]], device=dev), torch.tensor([len(node_seq)], device=dev))

    def append_walks(self, node_seqs, weights) -> None:
        dev = self.data.edge_index.device
        lengths = torch.tensor([len(seq) for seq in node_seqs], device=dev)
        # one vectorised id look-up over all walks (the reference maps walk by walk, path_data.py:139-142)
        flat = self.mapping.to_idxs(np.concatenate([np.asarray(seq) for seq in node_seqs]), device=dev)
        self.append_index_walks(flat, lengths, torch.tensor(weights, device=dev))

    def append_index_walks(self, flat_nodes: torch.Tensor, lengths: torch.Tensor, weights: torch.Tensor) -> None:
        """Walks given as one flat index tensor plus per-walk lengths (path_data.py:139-159 vectorised)."""
        dev = self.data.edge_index.device
        lengths = lengths.to(dev).long()
        if dev.type == "cuda":   # one scan + one kernel (csrc/walks.cu); the node offset of the container is folded in
            d = self.data
            chain = ops.walk_chain(lengths, base=int(d.num_nodes))
            d.edge_index = torch.cat([d.edge_index, chain], dim=1)
            d.node_sequence = torch.cat([d.node_sequence, flat_nodes.to(dev).long().unsqueeze(1)])
            d.dag_weight = torch.cat([d.dag_weight, weights.to(dev).to(d.dag_weight.dtype)])
            d.dag_num_edges = torch.cat([d.dag_num_edges, lengths - 1])
            d.dag_num_nodes = torch.cat([d.dag_num_nodes, lengths])
            d.num_nodes += chain.size(1) + lengths.numel()
            return
        # a container that lives in host memory (ingest before the upload): index plumbing on the host
        total = int(lengths.sum())
        pos = torch.arange(total, device=dev)
        chain = torch.stack([pos[:-1], pos[1:]])
        keep = torch.ones(chain.size(1), dtype=torch.bool, device=dev)
        ends = torch.cumsum(lengths, 0)
        keep[ends[:-1] - 1] = False  # drop the links between consecutive walks
        self._append_data(chain[:, keep], flat_nodes.to(dev).long().unsqueeze(1), weights.to(dev), lengths - 1, lengths)

    def get_walk(self, i: int) -> tuple:
        start = int(self.data.dag_num_nodes[:i].sum())
        end = start + int(self.data.dag_num_nodes[i])
        return tuple(self.mapping.to_ids(self.data.node_sequence[start:end].squeeze(1)).tolist())

    def map_node_seq(self, node_seq) -> list:
        return self.mapping.to_ids(node_seq).tolist()

    def __str__(self) -> str:
        return f"PathData with {self.num_paths} paths with total weight {self.data.dag_weight.sum().item()}"
