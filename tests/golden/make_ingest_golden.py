"""Golden vectors of the ingest step, produced by the REFERENCE'S OWN ``io/pandas.py`` (with its own
``core/index_map.py`` and ``core/path_data.py``), executed from /root/reference by ``oracle/ref_loader.io_module``.

    python tests/golden/make_ingest_golden.py        # needs /root/reference (read-only mount)

Event tables and n-gram files are seeded random; inputs and outputs are stored together in
``tests/golden/ingest_golden.npz``.  Events with equal time stamps are kept in input order (stable), see
``oracle/ingest.py``.
"""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ingest_golden.npz")
EVENT_CASES = [  # (nodes, events, horizon, string ids, multiedges, time_rescale)
    (6, 40, 10, True, False, 1),
    (50, 2000, 300, True, True, 1),
    (300, 20000, 5000, False, False, 7),
    (40, 3000, 50, False, True, 1),
]
PATH_CASES = [(5, 30, 6, True), (60, 2000, 9, True), (25, 500, 4, False)]  # (nodes, walks, max length, weight column)


def main() -> None:
    assert ref_loader.available(), "reference tree not mounted"
    io = ref_loader.io_module()
    rng = np.random.default_rng(20261019)
    out: dict[str, np.ndarray] = {}
    for i, (n, m, horizon, strings, multi, rescale) in enumerate(EVENT_CASES):
        v, w = rng.integers(0, n, m), rng.integers(0, n, m)
        if strings:
            v, w = np.array([f"n{x:03d}" for x in v]), np.array([f"n{x:03d}" for x in w])
        t = rng.integers(0, horizon, m)
        weight = rng.integers(1, 9, m).astype(np.float64)
        df = pd.DataFrame({"v": v, "w": w, "t": t, "weight": weight})
        g = io.df_to_temporal_graph(df.copy(), multiedges=multi, time_rescale=rescale)
        out[f"ev{i}_v"], out[f"ev{i}_w"], out[f"ev{i}_t"], out[f"ev{i}_weight"] = v, w, t, weight
        out[f"ev{i}_multiedges"], out[f"ev{i}_rescale"] = np.bool_(multi), np.int64(rescale)
        out[f"ev{i}_out_node_ids"] = np.asarray(g.mapping.node_ids).astype(str if strings else np.int64)
        out[f"ev{i}_out_edge_index"] = g.data.edge_index.numpy()
        out[f"ev{i}_out_time"] = g.data.time.numpy()
        out[f"ev{i}_out_weight"] = g.data.edge_weight.numpy()
        out[f"ev{i}_out_num_nodes"] = np.int64(g.data.num_nodes)
    tmp = tempfile.mkdtemp()
    for i, (n, p, max_len, weighted) in enumerate(PATH_CASES):
        lines = []
        for _ in range(p):
            walk = [f"s{x}" for x in rng.integers(0, n, rng.integers(1, max_len + 1))]
            lines.append(",".join(walk + ([str(float(rng.integers(1, 6)))] if weighted else [])))
        path = os.path.join(tmp, f"walks{i}.ngram")
        with open(path, "w") as f:
            f.write("\n".join(lines) + "\n")
        pdata = io.read_csv_path_data(path, weight=weighted)
        out[f"pa{i}_lines"], out[f"pa{i}_weighted"] = np.array(lines), np.bool_(weighted)
        out[f"pa{i}_out_node_ids"] = np.asarray(pdata.mapping.node_ids).astype(str)
        out[f"pa{i}_out_edge_index"] = pdata.data.edge_index.numpy()
        out[f"pa{i}_out_node_sequence"] = pdata.data.node_sequence.numpy()
        out[f"pa{i}_out_dag_weight"] = pdata.data.dag_weight.numpy()
        out[f"pa{i}_out_dag_num_nodes"] = pdata.data.dag_num_nodes.numpy()
        out[f"pa{i}_out_dag_num_edges"] = pdata.data.dag_num_edges.numpy()
    np.savez_compressed(OUT, **out)
    print(f"wrote {OUT}: {len(out)} arrays, {os.path.getsize(OUT) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
