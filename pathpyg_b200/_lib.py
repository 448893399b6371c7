"""ctypes binding of ``libpathpyg_b200.so`` (C ABI declared in ``include/pathpyg_b200.h``).

The library is built in-tree by ``pathpyg_b200/csrc/Makefile`` (nvcc, sm_100a).  There is no
CPU implementation behind these entry points: if the shared object is missing, or no CUDA
device is present, every compute call raises.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, c_char_p, c_double, c_int, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PATHPYG_B200_LIB", os.path.join(_HERE, "_C", "libpathpyg_b200.so"))  # override: instrumented builds
CSRC_DIR = os.path.join(_HERE, "csrc")

OK, ERR_INVALID, ERR_CUDA, ERR_WORKSPACE, ERR_EMPTY = 0, 1, 2, 3, 4

# element type codes (PPG_F32 ...)
F32, F64, I64, I32 = 0, 1, 2, 3
PAIR_RULES = {"src": 0, "dst": 1, "max": 2, "mul": 3, "add": 4}
REDUCTIONS = {"sum": 0, "add": 0, "mean": 1, "min": 2, "max": 3}
TIME_I64, TIME_F64, TIME_I64_F32DELTA = 0, 1, 2
TIME_GROUPED = 0x100

_p = c_void_p
_i64 = c_int64
_ph_i64 = POINTER(c_int64)
_ph_int = POINTER(c_int)

# name -> (restype, argtypes); must list every symbol of include/pathpyg_b200.h
PROTOTYPES = {
    "ppg_abi_version": (c_int, []),
    "ppg_last_error": (c_char_p, []),
    "ppg_launch_count": (ctypes.c_uint64, []),
    "ppg_profile_begin": (c_int, []),
    "ppg_profile_end": (c_int, [POINTER(ctypes.c_float), _ph_i64, _ph_int, _ph_int, c_int, _ph_int]),
    "ppg_result_read": (c_int, [_p, _ph_i64, _ph_int, _p]),
    "ppg_lift_order_workspace_bytes": (c_size_t, [_i64, _i64]),
    "ppg_lift_order_count": (c_int, [_p, _i64, _i64, _p, c_size_t, _ph_i64, _p]),
    "ppg_lift_order_fill": (c_int, [_p, _i64, _i64, _i64, _p, _p]),
    "ppg_pair_attributes": (c_int, [_p, _i64, _p, _i64, c_int, c_int, _p, _p, _p]),
    "ppg_lift_temporal_workspace_bytes": (c_size_t, [_i64, _i64]),
    "ppg_lift_temporal_group": (c_int, [_p, _i64, _i64, _p, c_size_t, _p]),
    "ppg_lift_temporal_count": (c_int, [_p, _p, _i64, _i64, c_int, _i64, c_double, _p, c_size_t, _ph_i64, _p]),
    "ppg_lift_temporal_fill": (c_int, [_p, _i64, _i64, _i64, _p, _p]),
    "ppg_lift_temporal_views": (c_int, [_p, _i64, _i64, POINTER(c_void_p)]),
    "ppg_chain_heavy_default": (c_int, []),
    "ppg_chain_tile_slots": (c_int, []),
    "ppg_chain_scan_workspace_bytes": (c_size_t, [_i64]),
    "ppg_chain_first_tiles": (c_int, [_p, _i64, _i64, _p, _p, _p, _p, c_int, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "ppg_chain_count_sorted": (c_int, [_p, _i64, _p, _p, _p, _i64, _p, _p, _p, c_size_t, _p, _p, _p, _p, _p, _p]),
    "ppg_chain_tiles": (c_int, [_i64, _i64, _i64, _p, _p, _p, _p, _p, _p, _p, _p, _p, c_int, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "ppg_chain_node_ptr": (c_int, [_p, _i64, _p, c_int, c_int, _p]),
    "ppg_chain_scan_nodes": (c_int, [_p, c_int, c_int, _i64, _p, c_size_t, _p, _p]),
    "ppg_chain_count_sorted_next": (c_int, [_p, _i64, _p, _p, c_int, c_int, _i64, _p, _p, _p, c_size_t, _p, _p, _p, _p]),
    "ppg_chain_heads": (c_int, [_p, _p, _p, _i64, _p, c_size_t, _p, _p, c_int, _p, _p, _p]),
    "ppg_chain_heavy_workspace_bytes": (c_size_t, [_i64, _i64, _i64]),
    "ppg_chain_heavy_fix": (c_int, [_p, _i64, _i64, _i64, _p, _p, _p, _p, _p, _p, _p, c_size_t, _p]),
    "ppg_chain_tiles_dist": (c_int, [_i64, _i64, _p, _p, _p, _p, c_int, _p, _p, _p, _p, _p, _p, c_int, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "ppg_chain_dest_bounds": (c_int, [_p, c_int, _i64, _p, c_int, _p, _p, _p]),
    "ppg_chain_pack": (c_int, [_p, _p, _p, _p, _i64, _p, _p]),
    "ppg_chain_unpack": (c_int, [_p, _i64, _p, _p, c_int, _p, _p, c_int, _p, _p, c_size_t, _p, _p, _p, _p, _p, _p]),
    "ppg_merge_sorted_tiles": (_i64, [_i64]),
    "ppg_chain_heavy_records_workspace_bytes": (c_size_t, [_i64, _i64, _i64]),
    "ppg_chain_heavy_fix_records": (c_int, [_p, _i64, _i64, _i64, _p, _p, _p, _p, _p, c_size_t, _p]),
    "ppg_merge_sorted": (c_int, [POINTER(c_void_p), _ph_i64, POINTER(c_void_p), c_int, _i64, _i64, _i64, _p, _p, _p, _p, _p, _p, _p, _p]),
    "ppg_merge_sorted_fill": (c_int, [_p, _p, _p, _p, _i64, _p, _p, _p, _p]),
    "ppg_chain_fill": (c_int, [_p, _p, _p, _p, _i64, _p, _p, _p]),
    "ppg_chain_widen": (c_int, [_p, c_int, _i64, _p, _p]),
    "ppg_rows_minmax_workspace_bytes": (c_size_t, [_i64]),
    "ppg_rows_minmax": (c_int, [_p, _i64, _i64, _p, c_size_t, _ph_i64, _ph_i64, _ph_int, _p]),
    "ppg_unique_rows_workspace_bytes": (c_size_t, [_i64, c_int]),
    "ppg_unique_rows_sort": (c_int, [_p, _i64, _i64, _ph_i64, _ph_int, c_int, _p, c_size_t, _p, _ph_i64, _p]),
    "ppg_unique_rows_gather": (c_int, [_p, _i64, _i64, _p, c_int, _i64, _p, _p]),
    "ppg_coalesce_workspace_bytes": (c_size_t, [_i64, _i64]),
    "ppg_coalesce_sort": (c_int, [_p, _i64, _p, _i64, _i64, _p, c_size_t, _p, _ph_i64, _p]),
    "ppg_extend_rows": (c_int, [_p, _i64, _i64, _p, _i64, _p, _p]),
    "ppg_coalesce_fill": (c_int, [_p, _i64, _i64, _i64, _p, c_int, c_int, _p, _p, _p]),
    "ppg_lift_limit": (c_int, [_p, c_int, _i64, _i64, _i64, _p]),
    "ppg_route_workspace_bytes": (c_size_t, [_i64]),
    "ppg_route_count": (c_int, [_p, _i64, _p, _p, c_int, _p, c_size_t, _p, _p]),
    "ppg_route_pack": (c_int, [_p, _i64, _p, _p, _i64, _p, c_int, _p, _p, POINTER(c_void_p), _p, _p, _p]),
    "ppg_route_unpack": (c_int, [_p, _i64, _p, _p, _p, _p, c_int, _p, _p]),
    "ppg_merge_records_workspace_bytes": (c_size_t, [_i64, _i64, _i64]),
    "ppg_merge_records_sort": (c_int, [_p, _i64, _i64, _i64, _i64, _p, c_size_t, _p, _p]),
    "ppg_merge_records_fill": (c_int, [_p, _p, _i64, _i64, _i64, _i64, _i64, _p, _p, _p, _p]),
    "ppg_extend_owned_rows": (c_int, [_p, _i64, _i64, _p, _p, _i64, _p, _p]),
    "ppg_sort_pairs_workspace_bytes": (c_size_t, [_i64, c_int]),
    "ppg_sort_pairs_u64": (c_int, [_p, _p, _i64, c_int, _p, c_size_t, POINTER(ctypes.c_float), _p]),
    "ppg_csc_workspace_bytes": (c_size_t, [_i64, _i64]),
    "ppg_csc_build": (c_int, [_p, _i64, _i64, _i64, _p, c_size_t, _p, _p, _p, _p]),
    "ppg_csc_build_async": (c_int, [_p, _i64, _i64, _i64, _p, c_size_t, _p, _p, _p, _p]),
    "ppg_gcn_norm": (c_int, [_p, _p, _p, _p, _i64, _i64, _p, _p, _p, _p, _p]),
    "ppg_colptr_counts": (c_int, [_p, _i64, _p, _p]),
    "ppg_spmm_csc": (c_int, [_p, _p, _p, _p, _p, _i64, _i64, _p, c_int, _p, _p]),
    "ppg_linear": (c_int, [_p, _p, _i64, _i64, _p, _p, _i64, _p, _p, _i64, c_int, _p, _p]),
    "ppg_gcn_fused_supported": (c_int, [_i64, _i64]),
    "ppg_gcn_layer_fused": (c_int, [_p, _p, _p, _p, _p, _p, _p, _i64, _i64, _i64, c_int, _p, _p]),
    "ppg_bipartite_fused": (c_int, [_p, _p, _p, _p, _p, _p, _p, _i64, _i64, _i64, c_int, _p, _p]),
    "ppg_counts_to_offsets_workspace_bytes": (c_size_t, [_i64]),
    "ppg_counts_to_offsets": (c_int, [_p, _i64, _i64, _p, c_size_t, _p, _p]),
    "ppg_expand_offsets": (c_int, [_p, _i64, _i64, _p, c_int, _p, _p, _p]),
    "ppg_walk_chain": (c_int, [_p, _i64, _i64, _i64, _p, _p]),
    "ppg_bincount": (c_int, [_p, _i64, _i64, _p, _p, _p]),
    "ppg_gcn_tc_supported": (c_int, [_i64, _i64]),
    "ppg_gcn_layer_tc": (c_int, [_p, _p, _p, _p, _p, _p, _p, _i64, _i64, _i64, _i64, c_int, _p, _p]),
    "ppg_act_backward_workspace_bytes": (c_size_t, [_i64, _i64]),
    "ppg_act_backward": (c_int, [_p, _p, _p, _i64, _i64, c_int, _p, _p, _p, _p, c_size_t, _p]),
    "ppg_atb_workspace_bytes": (c_size_t, [_i64, _i64, _i64]),
    "ppg_atb": (c_int, [_p, _p, _i64, _i64, _i64, _p, _p, c_size_t, _p]),
    "ppg_gather_f32": (c_int, [_p, _p, _i64, _p, _p]),
    "ppg_sorted_ids_ptr": (c_int, [_p, _i64, _i64, _p, _p]),
    "ppg_segment_sum_workspace_bytes": (c_size_t, [_i64]),
    "ppg_segment_sum": (c_int, [_p, _p, _p, _i64, _i64, _p, c_size_t, _p, _p]),
    "ppg_edge_ratio": (c_int, [_p, _p, _p, _i64, _p, _p]),
    "ppg_walk_counts_workspace_bytes": (c_size_t, [_i64, c_int]),
    "ppg_walk_counts": (c_int, [_p, _i64, _i64, c_int, _p, c_size_t, _ph_i64, _ph_i64, _p]),
    "ppg_temporal_paths_workspace_bytes": (c_size_t, [_i64, _i64, _i64]),
    "ppg_temporal_paths": (c_int, [_p, _i64, _i64, _p, _i64, _i64, _i64, _p, c_size_t, _p, _p, _ph_int, _p]),
    "ppg_temporal_closeness": (c_int, [_p, _i64, _p, _p]),
    "ppg_temporal_betweenness_workspace_bytes": (c_size_t, [_i64, _i64, _i64]),
    "ppg_temporal_betweenness": (c_int, [_p, _i64, _i64, _p, _i64, _p, _p, _p, _p, _p, _p, _p, _i64, _p, c_size_t, _p, _p]),
    "ppg_weighted_log_sum_workspace_bytes": (c_size_t, []),
    "ppg_weighted_log_sum": (c_int, [_p, _p, _p, _p, _i64, _i64, _i64, _p, c_size_t, POINTER(c_double), _p]),
}
ACT_NONE, ACT_ELU = 0, 1


class LibraryMissing(RuntimeError):
    pass


class EmptyLiftError(RuntimeError, ValueError):
    """No time-respecting pair.  The reference fails in ``torch.cat([])`` (temporal.py:53), which is a
    RuntimeError up to torch 2.8 (its pin) and a ValueError in newer releases -- this is both."""


_lib = None


def build(verbose: bool = False) -> str:
    """Compile the CUDA sources for sm_100a with nvcc (cross-compiles without a GPU)."""
    proc = subprocess.run(["make", "-j", str(min(8, os.cpu_count() or 1)), "-C", CSRC_DIR], capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("building libpathpyg_b200.so failed:\n" + proc.stdout + proc.stderr)
    if verbose:
        print(proc.stdout)
    return LIB_PATH


def load() -> ctypes.CDLL:
    """Load the shared object and bind every prototype.  Raises LibraryMissing if it was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise LibraryMissing(
            f"{LIB_PATH} not found: build it with `make -C {CSRC_DIR}` (or `python -c 'import __graft_entry__ as g; "
            "g.build()'`). pathpyg_b200 has no CPU fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.ppg_abi_version() != 1:
        raise RuntimeError(f"ABI version mismatch: library reports {lib.ppg_abi_version()}, binding expects 1")
    _lib = lib
    return lib


def check(rc: int) -> None:
    """Map a status code to the exception the reference would raise."""
    if rc == OK:
        return
    msg = load().ppg_last_error().decode("utf-8", "replace")
    if rc == ERR_INVALID:
        raise ValueError(msg)
    if rc == ERR_EMPTY:
        raise EmptyLiftError("torch.cat(): expected a non-empty list of Tensors (" + msg + ")")
    raise RuntimeError(msg)
