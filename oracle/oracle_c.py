"""TEST INFRASTRUCTURE — ctypes front end of oracle/oracle.c (see that file's header for the rules).

``build()`` compiles the C restatement with gcc into oracle/_build/liboracle.so (git-ignored, travels
to the GPU box with the snapshot).  Functions take/return numpy arrays.
"""
from __future__ import annotations

import ctypes as C
import pathlib
import subprocess
from typing import Optional, Sequence

import numpy as np

from .oracle_py import OutOfTable, _CELL_TYPES, last_column_mask

_HERE = pathlib.Path(__file__).resolve().parent
_SO = _HERE / "_build" / "liboracle.so"
_lib: Optional[C.CDLL] = None


def build(force: bool = False) -> pathlib.Path:
    src = _HERE / "oracle.c"
    if force or not _SO.exists() or _SO.stat().st_mtime < src.stat().st_mtime:
        _SO.parent.mkdir(exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-std=c11", "-o", str(_SO), str(src)])
    return _SO


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = C.CDLL(str(build()))
        _lib.oracle_result_count.restype = C.c_int64
        _lib.oracle_result_rows.restype = C.c_int64
        _lib.oracle_result_nodes.restype = C.c_int64
    return _lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def build_bit_table(weights: Sequence[int], max_mass: int, compression: int = 32) -> np.ndarray:
    """oracle_build_bit_table: literal loop of mass_table.py:207-248."""
    if compression not in _CELL_TYPES:
        raise ValueError(f"The compression rate {compression} is not compatible with the table setup.")
    w = np.ascontiguousarray(weights, dtype=np.int64)
    max_col = int(np.ceil((max_mass + 1) / compression))
    out = np.empty((len(w), max_col), dtype=_CELL_TYPES[compression])
    rc = lib().oracle_build_bit_table(_ptr(w), C.c_int(len(w)), C.c_int64(max_mass), C.c_int(compression),
                                      C.c_uint64(last_column_mask(max_mass, compression)), _ptr(out))
    if rc:
        raise RuntimeError(f"oracle_build_bit_table rc={rc}")
    return out


def is_valid(table: np.ndarray, compression: int, target: int, thr: int) -> bool:
    ok = C.c_int(0)
    rc = lib().oracle_is_valid(_ptr(table), C.c_int(table.shape[0]), C.c_int64(table.shape[1]), C.c_int(compression),
                               C.c_int64(target), C.c_int64(thr), C.byref(ok))
    if rc == 1:
        raise OutOfTable("value not in the DP table")
    return bool(ok.value)


def explain(table: np.ndarray, compression: int, weights, is_mod, ind, target: int, thr: int, max_mods,
            with_memo: bool = True):
    """-> (rows uint8[n_rows], offsets int64[n_sol+1], nodes).  max_mods None/inf = unbounded."""
    w = np.ascontiguousarray(weights, dtype=np.int64)
    im = np.ascontiguousarray(is_mod, dtype=np.uint8)
    iv = np.ascontiguousarray(ind, dtype=np.int64)
    mm = -1 if (max_mods is None or max_mods == float("inf")) else int(np.ceil(max(max_mods, 0)))
    res = C.c_void_p()
    rc = lib().oracle_explain(_ptr(table), C.c_int(table.shape[0]), C.c_int64(table.shape[1]), C.c_int(compression),
                              _ptr(w), _ptr(im), _ptr(iv), C.c_int64(target), C.c_int64(thr), C.c_int64(mm),
                              C.c_int(1 if with_memo else 0), C.byref(res))
    if rc == 1:
        raise OutOfTable("value not in the DP table")
    if rc:
        raise RuntimeError(f"oracle_explain rc={rc}")
    try:
        n = lib().oracle_result_count(res)
        nr = lib().oracle_result_rows(res)
        nodes = lib().oracle_result_nodes(res)
        rows = np.empty(max(nr, 1), dtype=np.uint8)
        off = np.empty(n + 1, dtype=np.int64)
        lib().oracle_result_fetch(res, _ptr(rows), _ptr(off))
    finally:
        lib().oracle_result_free(res)
    return rows[:nr], off, nodes


def length_bound(table: np.ndarray, compression: int, weights, is_mod, ind, target: int, thr: int, max_mods: int,
                 max_len: int, direction: str) -> int:
    if direction not in ("lower", "upper"):
        raise NotImplementedError(f"Support for '{direction}' is currently not given.")
    w = np.ascontiguousarray(weights, dtype=np.int64)
    im = np.ascontiguousarray(is_mod, dtype=np.uint8)
    iv = np.ascontiguousarray(ind, dtype=np.int64)
    out = C.c_int64(0)
    rc = lib().oracle_length_bound(_ptr(table), C.c_int(table.shape[0]), C.c_int64(table.shape[1]),
                                   C.c_int(compression), _ptr(w), _ptr(im), _ptr(iv), C.c_int64(target),
                                   C.c_int64(thr), C.c_int64(max_mods), C.c_int64(max_len),
                                   C.c_int(1 if direction == "lower" else 0), C.byref(out))
    if rc == 1:
        raise OutOfTable("value not in the DP table")
    return int(out.value)


def is_valid_batch(table: np.ndarray, compression: int, target, thr) -> np.ndarray:
    """oracle_is_valid over arrays -> uint8 codes: 0 not valid, 1 valid, 2 out of table."""
    tg = np.ascontiguousarray(target, dtype=np.int64)
    th = np.ascontiguousarray(thr, dtype=np.int64)
    out = np.empty(len(tg), dtype=np.uint8)
    lib().oracle_is_valid_batch(_ptr(table), C.c_int(table.shape[0]), C.c_int64(table.shape[1]), C.c_int(compression),
                                _ptr(tg), _ptr(th), C.c_int64(len(tg)), _ptr(out))
    return out


def explain_batch_keys(table: np.ndarray, compression: int, weights, is_mod, ind, target, thr, max_mods,
                       with_memo: bool = True):
    """oracle_explain over arrays -> (counts int64[n] (-1 = out of table), keys uint64[sum]): every composition as
    one little-endian 8-byte key (ascending row indices, zero padded), sorted ascending inside each call."""
    w = np.ascontiguousarray(weights, dtype=np.int64)
    im = np.ascontiguousarray(is_mod, dtype=np.uint8)
    iv = np.ascontiguousarray(ind, dtype=np.int64)
    tg = np.ascontiguousarray(target, dtype=np.int64)
    th = np.ascontiguousarray(thr, dtype=np.int64)
    mm = -1 if (max_mods is None or max_mods == float("inf")) else int(np.ceil(max(max_mods, 0)))
    counts = np.empty(len(tg), dtype=np.int64)
    cap = max(1024, 8 * len(tg))
    while True:
        keys = np.empty(cap, dtype=np.uint64)
        need = C.c_int64(0)
        rc = lib().oracle_explain_batch_keys(_ptr(table), C.c_int(table.shape[0]), C.c_int64(table.shape[1]),
                                             C.c_int(compression), _ptr(w), _ptr(im), _ptr(iv), _ptr(tg), _ptr(th),
                                             C.c_int64(len(tg)), C.c_int64(mm), C.c_int(1 if with_memo else 0),
                                             _ptr(counts), _ptr(keys), C.c_int64(cap), C.byref(need))
        if rc:
            raise RuntimeError(f"oracle_explain_batch_keys rc={rc}")
        if need.value <= cap:
            return counts, keys[: need.value]
        cap = int(need.value)
