"""The C-ABI library: loads without a GPU, exports every symbol include/pathpyg_b200.h declares,
and the binding lists exactly those.  No compute calls here."""
import os
import re

import pytest

from pathpyg_b200 import _lib

HEADER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "pathpyg_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ppg_[a-z0-9_]+)\s*\(", text)))


def test_library_loads_and_reports_abi_version():
    lib = _lib.load()
    assert lib.ppg_abi_version() == 1
    assert lib.ppg_last_error() is not None


def test_every_declared_symbol_is_exported_and_bound():
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(_lib.PROTOTYPES) == names, "binding and header disagree"


def test_workspace_queries_need_no_gpu():
    lib = _lib.load()
    assert lib.ppg_lift_order_workspace_bytes(1000, 100) > 1000 * 12
    assert lib.ppg_lift_temporal_workspace_bytes(1000, 100) > 1000 * 32
    assert lib.ppg_unique_rows_workspace_bytes(1000, 34) > 1000 * 24
    assert lib.ppg_coalesce_workspace_bytes(1000, 100) > 1000 * 24
    assert lib.ppg_csc_workspace_bytes(1000, 100) > 1000 * 16
    # monotone in the problem size
    assert lib.ppg_coalesce_workspace_bytes(2000, 100) > lib.ppg_coalesce_workspace_bytes(1000, 100)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.LibraryMissing):
        _lib.load()
