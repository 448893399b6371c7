"""Run the REFERENCE'S OWN test files against this package.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  ``python -m oracle.ref_suite <pytest args>`` makes
``import pathpyG`` resolve to ``pathpyg_b200`` (module by module), provides the three ``torch_geometric`` names the
reference's test files import (``Data``, ``EdgeIndex`` -> this package's stand-ins, ``get_random_edge_index``), and
hands over to pytest with the reference's test paths.  On a machine without a GPU every test that reaches a kernel
fails with the package's "needs a CUDA device" error; ``tests/test_reference_suite.py`` accepts those and nothing else.
"""
from __future__ import annotations

import importlib
import os
import sys
import tempfile
import types

SUBMODULES = ["core.graph", "core.index_map", "core.path_data", "core.temporal_graph", "core.multi_order_model",
              "algorithms.centrality", "algorithms.temporal", "algorithms.lift_order", "algorithms.shortest_paths",
              "algorithms.rolling_time_window", "algorithms.components", "algorithms.weisfeiler_leman", "io.pandas",
              "nn.dbgnn", "utils.dbgnn", "utils.convert"]


def install_alias() -> None:
    import torch

    import pathpyg_b200 as pp

    sys.modules["pathpyG"] = pp
    for sub in ("core", "algorithms", "io", "nn", "utils"):
        sys.modules[f"pathpyG.{sub}"] = getattr(pp, sub)
    for name in SUBMODULES:
        sys.modules[f"pathpyG.{name}"] = importlib.import_module(f"pathpyg_b200.{name}")

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    mod("torch_geometric", EdgeIndex=pp.EdgeIndex)
    mod("torch_geometric.data", Data=pp.Data)
    mod("torch_geometric.edge_index", EdgeIndex=pp.EdgeIndex)
    mod("torch_geometric.utils")
    mod("torch_geometric.testing",
        get_random_edge_index=lambda rows, cols, edges: torch.stack([torch.randint(0, rows, (edges,)), torch.randint(0, cols, (edges,))]))


def main(argv) -> int:
    import pytest

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    install_alias()
    scratch = tempfile.mkdtemp(prefix="ref_suite_")          # the reference tree is read-only: ini file and rootdir live here
    ini = os.path.join(scratch, "pytest.ini")
    open(ini, "w").write("[pytest]\n")
    return int(pytest.main(["-q", "-p", "no:cacheprovider", f"--rootdir={scratch}", "-c", ini, *argv]))


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
