"""Golden vectors of the model-selection statistics, produced by the REFERENCE'S OWN method bodies.

    python tests/golden/make_selection_golden.py        # needs /root/reference (read-only mount)

``oracle/ref_loader.selection_methods`` compiles ``MultiOrderModel.get_mon_dof``, the three log-likelihoods and
``likelihood_ratio_test`` (``src/pathpyG/core/multi_order_model.py:243-459``) and ``Graph.degrees`` /
``Graph.transition_probabilities`` (``src/pathpyG/core/graph.py:486-533``) from ``/root/reference`` unmodified and runs
them on seeded random walks.  Inputs (flat walks, lengths, weights) and outputs are stored together in
``tests/golden/selection_golden.npz`` so that the GPU box (no /root/reference) can check the oracle and the CUDA
path against them.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import mom, ref_loader  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "selection_golden.npz")
CASES = [  # (nodes, walks, max length, max weight, max_order)
    (8, 40, 5, 3, 3),
    (30, 500, 7, 5, 3),
    (50, 3000, 9, 4, 4),
    (200, 4000, 6, 1, 3),
]


def random_walks(gen, n, p, max_len, max_w):
    lengths = torch.randint(1, max_len + 1, (p,), generator=gen)
    flat = torch.randint(0, n, (int(lengths.sum()),), generator=gen)
    flat[:n] = torch.arange(n)  # every node occurs (multi_order_model.py:336-337)
    weights = torch.randint(1, max_w + 1, (p,), generator=gen).float()
    return flat, lengths, weights


def split(flat, lengths):
    out, o = [], 0
    for length in lengths.tolist():
        out.append(flat[o:o + length].tolist())
        o += length
    return out


def main() -> None:
    assert ref_loader.available(), "reference tree not mounted"
    Model, _ = ref_loader.selection_methods()
    gen = torch.Generator().manual_seed(20261018)
    out: dict[str, np.ndarray] = {}
    for i, (n, p, max_len, max_w, K) in enumerate(CASES):
        flat, lengths, weights = random_walks(gen, n, p, max_len, max_w)
        walks = mom.append_walks(split(flat, lengths), weights.tolist())
        layers = mom.from_path_data(walks, max_order=K)
        model, dag = Model(layers), ref_loader.ref_walks_data(walks)
        out[f"sel{i}_flat"], out[f"sel{i}_lengths"], out[f"sel{i}_weights"] = flat.numpy(), lengths.numpy(), weights.numpy()
        out[f"sel{i}_num_nodes"], out[f"sel{i}_max_order"] = np.int64(n), np.int64(K)
        out[f"sel{i}_dof_paths"] = np.array([model.get_mon_dof(k, "paths") for k in range(K + 1)], dtype=np.int64)
        out[f"sel{i}_dof_ngrams"] = np.array([model.get_mon_dof(k, "ngrams") for k in range(K + 1)], dtype=np.float64)
        out[f"sel{i}_llh"] = np.array([model.get_mon_log_likelihood(dag, k) for k in range(K + 1)], dtype=np.float64)
        out[f"sel{i}_llh_mid"] = np.array([model.get_intermediate_order_log_likelihood(dag, k) for k in range(1, K)], dtype=np.float64)
        tests = [model.likelihood_ratio_test(dag, k - 1, k) for k in range(1, K + 1)]
        out[f"sel{i}_lrt_reject"] = np.array([bool(t[0]) for t in tests])
        out[f"sel{i}_lrt_p"] = np.array([float(t[1]) for t in tests], dtype=np.float64)
        for k in range(1, K + 1):
            out[f"sel{i}_tp{k}"] = model.layers[k].transition_probabilities(edge_attr="edge_weight").numpy()
            out[f"sel{i}_tp_unit{k}"] = model.layers[k].transition_probabilities().numpy()
            out[f"sel{i}_indeg{k}"] = model.layers[k].degrees("in", "edge_weight", True).numpy()
            out[f"sel{i}_outdeg_unit{k}"] = model.layers[k].degrees("out", None, True).numpy()
    np.savez_compressed(OUT, **out)
    print(f"wrote {OUT}: {len(out)} arrays, {os.path.getsize(OUT) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
