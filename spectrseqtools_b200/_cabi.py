"""ctypes binding of libsst_b200.so (include/sst_b200.h) plus a small object layer on top.

There is NO CPU fallback: if the shared library is missing, or the machine has no sm_100 GPU, every
entry point raises ``DeviceUnavailable`` with the reason.  ``load()`` alone (symbol check) works
without a GPU so that CPU-only CI can verify the ABI.
"""
from __future__ import annotations

import ctypes as C
import os
import pathlib
import threading
import weakref
from typing import Dict, Optional, Sequence, Tuple

import numpy as np

_HERE = pathlib.Path(__file__).resolve().parent
LIB_PATH = _HERE / "libsst_b200.so"

# error codes (include/sst_b200.h)
SST_OK, SST_ERR_CUDA, SST_ERR_NO_DEVICE, SST_ERR_BAD_ARG, SST_ERR_COMPRESSION = 0, 1, 2, 3, 4
SST_ERR_TOO_MANY_ROWS, SST_ERR_TOO_DEEP, SST_ERR_NOMEM, SST_ERR_MEMO_FULL, SST_ERR_STATE = 5, 6, 7, 8, 9
SST_ERR_OUT_OF_TABLE = 10
SST_ERR_NAN, SST_ERR_INF = 11, 12
MODE_FREE, MODE_EXACT, MODE_MEMO = 0, 1, 2
STATUS_ZERO_IN_WINDOW, STATUS_OUT_OF_TABLE = 1, 2
VALID_NO, VALID_YES, VALID_OUT_OF_TABLE = 0, 1, 2
CLASS_VALID, CLASS_OUT_OF_TABLE, CLASS_SINGLETON = 1, 2, 4
BUDGET_INF = 1 << 30
KERNEL_SLOTS = ["build", "transpose", "is_valid", "phase_a", "explain_pass", "classify", "length_bound", "count_table"]

EXPORTS = [
    "sst_ctx_create", "sst_ctx_destroy", "sst_last_error", "sst_device_info", "sst_host_alloc", "sst_host_free",
    "sst_timer_start", "sst_timer_stop", "sst_timer_stop_at_run", "sst_stats_reset", "sst_kernel_ms", "sst_flush_l2", "sst_set_item_limit",
    "sst_table_build", "sst_table_upload", "sst_table_rebuild", "sst_table_info", "sst_table_download",
    "sst_table_download_masks", "sst_table_destroy", "sst_is_valid", "sst_valid_stage", "sst_valid_run",
    "sst_valid_fetch", "sst_valid_stage_f64", "sst_explain", "sst_explain_stage", "sst_explain_stage_f64", "sst_explain_stage_f64_uniform", "sst_explain_rec_width", "sst_explain_phase_ns",
    "sst_explain_run", "sst_explain_fetch", "sst_classify", "sst_classify_stage", "sst_classify_run", "sst_classify_fetch", "sst_classify_launch", "sst_classify_async", "sst_classify_wait", "sst_length_bounds",
    "sst_set_pass", "sst_last_pass", "sst_explain_cta_ns", "sst_explain_submit_f64", "sst_explain_collect", "sst_classify_async_packed", "sst_host_profile", "sst_trace_ms", "sst_explain_block_layout", "sst_explain_d2h_bytes",
    "sst_ladder_stage", "sst_ladder_round", "sst_ladder_revalidate", "sst_ladder_fetch", "sst_host_register", "sst_host_unregister", "sst_count_compositions_f64", "sst_is_valid_f64", "sst_set_record_split", "sst_explain_rec_layout",
]


class DeviceUnavailable(RuntimeError):
    """The CUDA library or a B200-class GPU is missing; the product path has no CPU fallback."""


class TableTooSmall(NotImplementedError):
    """A probed mass lies beyond the table (the reference raises NotImplementedError here)."""


_lib: Optional[C.CDLL] = None
_lock = threading.Lock()


def load() -> C.CDLL:
    """dlopen the library and declare signatures.  Needs no GPU."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not LIB_PATH.exists():
            raise DeviceUnavailable(
                f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a). "
                "spectrseqtools_b200 has no CPU fallback.")
        lib = C.CDLL(str(LIB_PATH))
        vp, i64p, u64p, u8p, i32p, u32p, fp = C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p
        sig = {
            "sst_ctx_create": (C.c_int, [C.c_int, C.POINTER(vp)]),
            "sst_ctx_destroy": (None, [vp]),
            "sst_last_error": (C.c_char_p, [vp]),
            "sst_device_info": (C.c_int, [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
            "sst_host_alloc": (vp, [vp, C.c_size_t]),
            "sst_host_free": (None, [vp, vp]),
            "sst_timer_start": (C.c_int, [vp]),
            "sst_timer_stop": (C.c_int, [vp, C.POINTER(C.c_float)]),
            "sst_timer_stop_at_run": (C.c_int, [vp, C.POINTER(C.c_float)]),
            "sst_stats_reset": (C.c_int, [vp]),
            "sst_kernel_ms": (C.c_int, [vp, fp, u64p]),
            "sst_flush_l2": (C.c_int, [vp, C.c_size_t]),
            "sst_set_item_limit": (C.c_int, [vp, C.c_uint64]),
            "sst_table_build": (C.c_int, [vp, i64p, C.c_int, C.c_int64, C.c_int, C.c_uint64, C.c_int, C.POINTER(vp)]),
            "sst_table_upload": (C.c_int, [vp, u64p, i64p, C.c_int, C.c_int64, C.POINTER(vp)]),
            "sst_table_rebuild": (C.c_int, [vp, vp]),
            "sst_table_info": (C.c_int, [vp, C.POINTER(C.c_int), C.POINTER(C.c_int64), C.POINTER(C.c_float), C.POINTER(C.c_float)]),
            "sst_table_download": (C.c_int, [vp, vp, u64p]),
            "sst_table_download_masks": (C.c_int, [vp, vp, C.c_int64, C.c_int64, u32p]),
            "sst_table_destroy": (None, [vp, vp]),
            "sst_is_valid": (C.c_int, [vp, vp, i64p, i64p, C.c_int64, u8p]),
            "sst_valid_stage": (C.c_int, [vp, i64p, i64p, C.c_int64]),
            "sst_valid_stage_f64": (C.c_int, [vp, fp, fp, C.c_int64, C.c_double, C.c_double]),
            "sst_explain_stage_f64": (C.c_int, [vp, vp, fp, fp, i32p, C.c_int64, i32p, u8p, C.c_double, C.c_double, C.c_int]),
            "sst_explain_stage_f64_uniform": (C.c_int, [vp, vp, fp, fp, C.c_int32, C.c_int64, i32p, u8p, C.c_double, C.c_double, C.c_int]),
            "sst_explain_rec_width": (C.c_int, [vp]),
            "sst_explain_phase_ns": (C.c_int, [vp, u64p]),
            "sst_valid_run": (C.c_int, [vp, vp]),
            "sst_valid_fetch": (C.c_int, [vp, u8p]),
            "sst_explain": (C.c_int, [vp, vp, i64p, i64p, i32p, u8p, C.c_int64, i32p, u8p, C.c_int, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
            "sst_explain_stage": (C.c_int, [vp, vp, i64p, i64p, i32p, u8p, C.c_int64, i32p, u8p]),
            "sst_explain_run": (C.c_int, [vp, vp, C.c_int, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
            "sst_explain_fetch": (C.c_int, [vp, u8p, u64p, u8p]),
            "sst_classify": (C.c_int, [vp, vp, fp, C.c_int64, fp, C.c_int, C.c_double, C.c_double, u8p]),
            "sst_classify_stage": (C.c_int, [vp, fp, C.c_int64, fp, C.c_int]),
            "sst_classify_run": (C.c_int, [vp, vp, C.c_double, C.c_double]),
            "sst_classify_fetch": (C.c_int, [vp, u8p]),
            "sst_classify_launch": (C.c_int, [vp, vp, C.c_double, C.c_double]),
            "sst_classify_async": (C.c_int, [vp, vp, fp, C.c_int64, fp, C.c_int, C.c_double, C.c_double, u8p]),
            "sst_classify_async_packed": (C.c_int, [vp, vp, fp, C.c_int64, fp, C.c_int, C.c_double, C.c_double, u8p]),
            "sst_classify_wait": (C.c_int, [vp]),
            "sst_explain_submit_f64": (C.c_int, [vp, vp, fp, fp, C.c_int32, C.c_int64, i32p, u8p, C.c_double, C.c_double, C.c_int, u8p, C.c_uint64]),
            "sst_explain_block_layout": (C.c_int, [C.c_int64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
            "sst_explain_collect": (C.c_int, [vp, vp, C.POINTER(C.c_uint64), C.POINTER(C.c_int)]),
            "sst_host_profile": (C.c_int, [C.c_int, u64p, u64p]),
            "sst_trace_ms": (C.c_int, [vp, C.c_int, fp]),
            "sst_count_compositions_f64": (C.c_int, [vp, vp, fp, fp, C.c_int64, C.c_double, C.c_double, u64p]),
            "sst_is_valid_f64": (C.c_int, [vp, vp, fp, fp, C.c_int64, C.c_double, C.c_double, u8p]),
            "sst_set_record_split": (C.c_int, [vp, C.c_int]),
            "sst_explain_rec_layout": (C.c_int, [vp, C.POINTER(C.c_int), C.POINTER(C.c_uint64), C.POINTER(C.c_int)]),
            "sst_host_register": (C.c_int, [vp, vp, C.c_size_t]),
            "sst_host_unregister": (C.c_int, [vp, vp]),
            "sst_ladder_stage": (C.c_int, [vp, fp, fp, u8p, C.c_int64]),
            "sst_ladder_round": (C.c_int, [vp, vp, C.c_double, C.c_double, C.c_double, C.c_int32, i32p, u8p, C.c_int, u32p,
                                           C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
            "sst_ladder_revalidate": (C.c_int, [vp, vp, C.c_double, C.c_double, C.POINTER(C.c_int64)]),
            "sst_ladder_fetch": (C.c_int, [vp, u8p, fp, fp, u8p]),
            "sst_explain_d2h_bytes": (C.c_uint64, [vp]),
            "sst_set_pass": (C.c_int, [vp, C.c_int]),
            "sst_last_pass": (C.c_int, [vp]),
            "sst_explain_cta_ns": (C.c_int, [vp, C.c_int, u64p, C.c_int, C.POINTER(C.c_int)]),
            "sst_length_bounds": (C.c_int, [vp, vp, C.c_int64, C.c_int64, C.c_int32, C.c_int32, i32p, u8p, C.c_uint64,
                                            C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
        }
        for name in EXPORTS:
            fn = getattr(lib, name)  # AttributeError here = ABI drift
            fn.restype, fn.argtypes = sig[name]
        _lib = lib
        return lib


_PTRS: Dict[int, tuple] = {}  # id(array) -> (weak reference, address): see _p


def _p(a: Optional[np.ndarray]):
    """Address of an array's data for a ``c_void_p`` argument.  ``ndarray.ctypes`` costs 2.5-4 us per use and a queued
    call passes ten of them, so the address of an array object that has been seen before is remembered (a weak reference
    guards against a recycled ``id``; arrays are not resized in place anywhere in this package)."""
    if a is None:
        return None
    k = id(a)
    hit = _PTRS.get(k)
    if hit is not None and hit[0]() is a:
        return hit[1]
    ptr = a.ctypes.data
    if len(_PTRS) > 256:
        _PTRS.clear()
    _PTRS[k] = (weakref.ref(a), ptr)
    return ptr


def _arr(x, dtype) -> np.ndarray:
    if type(x) is np.ndarray and x.dtype == dtype and x.flags.c_contiguous:
        return x
    return np.ascontiguousarray(x, dtype=dtype)


class Context:
    """One CUDA stream on one device + scratch.  Not thread-safe."""

    def __init__(self, device: int = 0):
        lib = load()
        h = C.c_void_p()
        rc = lib.sst_ctx_create(int(device), C.byref(h))
        if rc != SST_OK:
            raise DeviceUnavailable(
                f"no usable sm_100 (B200) GPU at index {device} (sst_ctx_create rc={rc}); "
                "spectrseqtools_b200 has no CPU fallback.")
        self._lib, self._h, self.device = lib, h, int(device)
        lib.sst_set_record_split(h, 1)  # queued batches bring their records back as 4 + k byte planes (SplitRecords)
        self._finalizer = weakref.finalize(self, lib.sst_ctx_destroy, h)

    # -- plumbing
    def _check(self, rc: int):
        if rc == SST_OK:
            return
        msg = (self._lib.sst_last_error(self._h) or b"").decode(errors="replace")
        if rc in (SST_ERR_BAD_ARG, SST_ERR_COMPRESSION, SST_ERR_TOO_MANY_ROWS):
            raise ValueError(msg)
        if rc == SST_ERR_NAN:  # the reference's int(round(nan))
            raise ValueError(msg)
        if rc == SST_ERR_INF:  # int(round(inf)) / int(np.ceil(inf))
            raise OverflowError(msg)
        if rc == SST_ERR_TOO_DEEP:
            raise NotImplementedError(msg)
        if rc == SST_ERR_NOMEM:
            raise MemoryError(msg)
        if rc == SST_ERR_MEMO_FULL:
            raise MemoFull(msg)
        if rc == SST_ERR_OUT_OF_TABLE:
            raise TableTooSmall(msg)
        raise RuntimeError(f"libsst_b200 rc={rc}: {msg}")

    def device_info(self) -> dict:
        sm, ma, mi = C.c_int(), C.c_int(), C.c_int()
        fr, to = C.c_uint64(), C.c_uint64()
        self._check(self._lib.sst_device_info(self._h, C.byref(sm), C.byref(ma), C.byref(mi), C.byref(fr), C.byref(to)))
        return dict(sm_count=sm.value, cc=(ma.value, mi.value), free_bytes=fr.value, total_bytes=to.value)

    def pinned_empty(self, shape, dtype) -> np.ndarray:
        """numpy array on page-locked host memory (freed when the array is garbage collected)."""
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        ptr = self._lib.sst_host_alloc(self._h, max(n, 1))
        if not ptr:
            raise MemoryError(f"cudaHostAlloc({n}) failed")
        buf = (C.c_uint8 * max(n, 1)).from_address(ptr)
        arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        weakref.finalize(buf, self._lib.sst_host_free, self._h, C.c_void_p(ptr))
        return arr

    def timer_start(self):
        self._check(self._lib.sst_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = C.c_float()
        self._check(self._lib.sst_timer_stop(self._h, C.byref(ms)))
        return float(ms.value)

    def timer_stop_at_run(self) -> float:
        """Device time from ``timer_start`` to the end of the device work of the last ``explain_run``."""
        ms = C.c_float()
        self._check(self._lib.sst_timer_stop_at_run(self._h, C.byref(ms)))
        return float(ms.value)

    def stats_reset(self):
        self._check(self._lib.sst_stats_reset(self._h))

    def kernel_stats(self) -> Dict[str, Tuple[float, int]]:
        ms = np.zeros(len(KERNEL_SLOTS), dtype=np.float32)
        n = np.zeros(len(KERNEL_SLOTS), dtype=np.uint64)
        self._check(self._lib.sst_kernel_ms(self._h, _p(ms), _p(n)))
        return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(KERNEL_SLOTS)}

    def set_item_limit(self, limit: int):
        """Blow-up guard: most partial compositions one level of a pass may hold (0 = device memory)."""
        self._check(self._lib.sst_set_item_limit(self._h, C.c_uint64(limit)))

    def set_pass(self, which: int):
        """0 = automatic (item pass for light batches, level-synchronous for heavy or deep ones), 1 = level-synchronous
        pass, 2 = depth-first item pass, 3 = direct pass (count table) (sst_set_pass)."""
        self._check(self._lib.sst_set_pass(self._h, int(which)))

    def last_pass(self) -> int:
        return int(self._lib.sst_last_pass(self._h))

    def cta_timestamps(self, enable: bool = True) -> np.ndarray:
        """Switch the per-CTA phase timestamps of the depth-first pass on / off and return those of the last recorded
        run as uint64[n_ctas, 8] (see sst_explain_cta_ns)."""
        out = np.zeros((1024, 8), dtype=np.uint64)
        n = C.c_int()
        self._check(self._lib.sst_explain_cta_ns(self._h, 1 if enable else 0, _p(out), 1024, C.byref(n)))
        return out[: n.value]

    def flush_l2(self, nbytes: int = 256 << 20):
        self._check(self._lib.sst_flush_l2(self._h, int(nbytes)))

    # -- tables
    def build_table(self, weights: Sequence[int], max_mass: int, compression: int, last_col_mask: int,
                    with_masks: bool = True) -> "DeviceTable":
        w = _arr(weights, np.int64)
        h = C.c_void_p()
        self._check(self._lib.sst_table_build(self._h, _p(w), len(w), int(max_mass), int(compression),
                                              C.c_uint64(last_col_mask), 1 if with_masks else 0, C.byref(h)))
        return DeviceTable(self, h, w)

    def upload_table(self, table: np.ndarray, weights: Sequence[int]) -> "DeviceTable":
        if table.dtype != np.uint64 or table.ndim != 2:
            raise ValueError("only uint64 tables (32 masses per cell) can be adopted by the device path")
        t = np.ascontiguousarray(table)
        w = _arr(weights, np.int64)
        if len(w) != t.shape[0]:
            raise ValueError("one weight per table row expected")
        h = C.c_void_p()
        self._check(self._lib.sst_table_upload(self._h, _p(t), _p(w), t.shape[0], t.shape[1], C.byref(h)))
        return DeviceTable(self, h, w)

    # -- hot path
    def is_valid(self, table: "DeviceTable", target, thr) -> np.ndarray:
        t, h = _arr(target, np.int64), _arr(thr, np.int64)
        out = np.empty(len(t), dtype=np.uint8)
        self._check(self._lib.sst_is_valid(self._h, table._h, _p(t), _p(h), len(t), _p(out)))
        return out

    def valid_stage(self, target, thr):
        t, h = _arr(target, np.int64), _arr(thr, np.int64)
        self._check(self._lib.sst_valid_stage(self._h, _p(t), _p(h), len(t)))
        self._staged_VP = len(t)

    def valid_stage_f64(self, mass: np.ndarray, thr: Optional[np.ndarray], precision: float, tolerance: float):
        m = _arr(mass, np.float64)
        h = None if thr is None else _arr(thr, np.float64)
        if h is not None and len(h) != len(m):
            raise ValueError("per-probe arrays differ in length")
        self._check(self._lib.sst_valid_stage_f64(self._h, _p(m), _p(h), len(m), float(precision), float(tolerance)))
        self._staged_VP = len(m)

    def is_valid_f64(self, table: "DeviceTable", mass: np.ndarray, thr: Optional[np.ndarray], precision: float, tolerance: float) -> np.ndarray:
        """Validity codes of a (small) batch with one synchronisation (sst_is_valid_f64)."""
        m = _arr(mass, np.float64)
        h = None if thr is None else _arr(thr, np.float64)
        if h is not None and len(h) != len(m):
            raise ValueError("per-probe arrays differ in length")
        out = np.empty(len(m), dtype=np.uint8)
        self._check(self._lib.sst_is_valid_f64(self._h, table._h, _p(m), _p(h), len(m), float(precision), float(tolerance), _p(out)))
        return out

    def explain_stage_f64(self, table: "DeviceTable", mass, thr, max_mods, ind, is_mod, precision, tolerance, with_memo):
        """``max_mods``: one int for the whole batch, or one int32 per peak."""
        m = _arr(mass, np.float64)
        h = None if thr is None else _arr(thr, np.float64)
        iv, im = _arr(ind, np.int32), _arr(is_mod, np.uint8)
        if h is not None and len(h) != len(m):
            raise ValueError("per-peak arrays differ in length")
        if len(iv) != table.R or len(im) != table.R:
            raise ValueError("ind / is_mod need one entry per table row")
        if np.ndim(max_mods) == 0:
            self._check(self._lib.sst_explain_stage_f64_uniform(self._h, table._h, _p(m), _p(h), int(max_mods), len(m), _p(iv), _p(im),
                                                                float(precision), float(tolerance), 1 if with_memo else 0))
        else:
            mm = _arr(max_mods, np.int32)
            if len(mm) != len(m):
                raise ValueError("per-peak arrays differ in length")
            self._check(self._lib.sst_explain_stage_f64(self._h, table._h, _p(m), _p(h), _p(mm), len(m), _p(iv), _p(im),
                                                        float(precision), float(tolerance), 1 if with_memo else 0))
        self._staged_P = len(m)

    def valid_run(self, table: "DeviceTable"):
        self._check(self._lib.sst_valid_run(self._h, table._h))

    def valid_fetch(self) -> np.ndarray:
        out = np.empty(self._staged_VP, dtype=np.uint8)
        self._check(self._lib.sst_valid_fetch(self._h, _p(out)))
        return out

    def classify_stage(self, observed: np.ndarray, offsets: np.ndarray):
        o, b = _arr(observed, np.float64), _arr(offsets, np.float64)
        self._check(self._lib.sst_classify_stage(self._h, _p(o), len(o), _p(b), len(b)))
        self._staged_C = (len(b), len(o))

    def classify_run(self, table: "DeviceTable", precision: float, tolerance: float):
        self._check(self._lib.sst_classify_run(self._h, table._h, float(precision), float(tolerance)))

    def classify_launch(self, table: "DeviceTable", precision: float, tolerance: float):
        """Queue the classification kernel without waiting (the next synchronous call on the context completes it)."""
        self._check(self._lib.sst_classify_launch(self._h, table._h, float(precision), float(tolerance)))

    def classify_async(self, table: "DeviceTable", observed: np.ndarray, offsets: np.ndarray, precision: float, tolerance: float,
                       packed: bool = False, out: Optional[np.ndarray] = None):
        """Whole classification on the side stream, no waiting; returns the pinned uint8[B, F] buffer the flags will
        land in — valid after ``classify_wait()`` and until the next classification on this context.  ``packed``: two
        flags per byte, uint8[B, ceil(F / 2)] (fragment f in the low nibble when f is even)."""
        o, b = _arr(observed, np.float64), _arr(offsets, np.float64)
        self._async_keep = (o, b)  # the copies are asynchronous: keep the host arrays alive
        B, F = len(b), len(o)
        if packed:
            half = (F + 1) // 2
            base = self._pinned("classify", B * half) if out is None else out
            if base.size < B * half:
                raise ValueError("out is too small for the packed flags")
            buf = base[: B * half]
            self._check(self._lib.sst_classify_async_packed(self._h, table._h, _p(o), F, _p(b), B, float(precision), float(tolerance), _p(base)))
            return buf.reshape(B, half)
        buf = self._pinned("classify", B * F)[: B * F]
        self._check(self._lib.sst_classify_async(self._h, table._h, _p(o), F, _p(b), B, float(precision), float(tolerance), _p(buf)))
        return buf.reshape(B, F)

    def classify_wait(self):
        self._check(self._lib.sst_classify_wait(self._h))
        self._async_keep = None

    def classify_fetch(self, copy: bool = True) -> np.ndarray:
        """-> uint8[B, F] of CLASS_* bits (breakage-major)."""
        B, F = self._staged_C
        buf = self._pinned("classify", B * F)[: B * F]
        self._check(self._lib.sst_classify_fetch(self._h, _p(buf)))
        out = buf.reshape(B, F)
        return out.copy() if copy else out

    def length_bounds(self, table: "DeviceTable", target: int, thr: int, max_mods: int, max_len: int, ind, is_mod,
                      memo_capacity: int = 0) -> Tuple[int, int]:
        """(lower, upper) bound on the number of nucleotides explaining the window (sst_length_bounds)."""
        iv, im = _arr(ind, np.int32), _arr(is_mod, np.uint8)
        lo, up = C.c_int64(), C.c_int64()
        self._check(self._lib.sst_length_bounds(self._h, table._h, int(target), int(thr), int(max_mods), int(max_len), _p(iv), _p(im),
                                                C.c_uint64(memo_capacity), C.byref(lo), C.byref(up)))
        return int(lo.value), int(up.value)

    def explain_stage(self, table: "DeviceTable", target, thr, max_mods, mode, ind, is_mod):
        t, h = _arr(target, np.int64), _arr(thr, np.int64)
        mm, mo = _arr(max_mods, np.int32), _arr(mode, np.uint8)
        iv, im = _arr(ind, np.int32), _arr(is_mod, np.uint8)
        if not (len(t) == len(h) == len(mm) == len(mo)):
            raise ValueError("per-peak arrays differ in length")
        if len(iv) != table.R or len(im) != table.R:
            raise ValueError("ind / is_mod need one entry per table row")
        self._check(self._lib.sst_explain_stage(self._h, table._h, _p(t), _p(h), _p(mm), _p(mo), len(t), _p(iv), _p(im)))
        self._staged_P = len(t)

    def explain_run(self, table: "DeviceTable", rec_width: int = 0, memo_capacity: int = 0) -> Tuple[int, int]:
        """rec_width 0 = automatic (smallest multiple of 8 holding the longest possible composition)."""
        nr, nc = C.c_uint64(), C.c_uint64()
        self._check(self._lib.sst_explain_run(self._h, table._h, int(rec_width), C.c_uint64(memo_capacity), C.byref(nr), C.byref(nc)))
        self._last = (int(nr.value), int(nc.value), int(self._lib.sst_explain_rec_width(self._h)))
        return int(nr.value), int(nc.value)

    def count_compositions_f64(self, table: "DeviceTable", mass, thr, precision: float, tolerance: float) -> np.ndarray:
        """uint64[P]: compositions per call, looked up (sst_count_compositions_f64); 2**64 - 1 = not known."""
        m = _arr(mass, np.float64)
        h = None if thr is None else _arr(thr, np.float64)
        if h is not None and len(h) != len(m):
            raise ValueError("per-call arrays differ in length")
        out = np.zeros(len(m), dtype=np.uint64)
        self._check(self._lib.sst_count_compositions_f64(self._h, table._h, _p(m), _p(h), len(m), float(precision), float(tolerance), _p(out)))
        return out

    def host_register(self, arr: np.ndarray):
        """Page-lock a caller-owned buffer (sst_host_register); ``host_unregister`` before it is freed."""
        self._check(self._lib.sst_host_register(self._h, _p(arr), arr.nbytes))

    def host_unregister(self, arr: np.ndarray):
        self._check(self._lib.sst_host_unregister(self._h, _p(arr)))

    def explain_submit_f64(self, table: "DeviceTable", mass, thr, max_mods: int, ind, is_mod, precision, tolerance, with_memo, out_block=None):
        """Queue a whole enumeration call (inputs in, staging, pass, results out) without waiting; ``explain_collect``
        completes it.  Results land in this context's pinned result block, or in ``out_block`` (uint8, page-locked by
        the caller, e.g. a registered shared-memory region)."""
        m = _arr(mass, np.float64)
        h = None if thr is None else _arr(thr, np.float64)
        iv, im = _arr(ind, np.int32), _arr(is_mod, np.uint8)
        if h is not None and len(h) != len(m):
            raise ValueError("per-peak arrays differ in length")
        if len(iv) != table.R or len(im) != table.R:
            raise ValueError("ind / is_mod need one entry per table row")
        P = len(m)
        lay = _BLOCK_LAYOUTS.get(P)
        if lay is None:
            so, oo, ro = C.c_uint64(), C.c_uint64(), C.c_uint64()
            self._check(self._lib.sst_explain_block_layout(P, C.byref(so), C.byref(oo), C.byref(ro)))
            if len(_BLOCK_LAYOUTS) > 64:
                _BLOCK_LAYOUTS.clear()
            lay = _BLOCK_LAYOUTS[P] = (so, oo, ro)
        so, oo, ro = lay
        if out_block is not None:
            block = out_block
            if block.dtype != np.uint8 or block.ndim != 1 or block.size <= ro.value:
                raise ValueError("out_block must be a flat uint8 array larger than the block's fixed part")
        else:
            block = self._pinned("block", ro.value + max(self.__dict__.get("_recs_hint", 0), 64 * P, 1 << 20))  # one block: one copy brings it back
        status = block[so.value: so.value + P]
        off = block[oo.value: oo.value + 4 * (P + 1)].view(np.uint32)
        recs = block[ro.value:]
        self._submitted = (m, h, iv, im, P, status, off, recs, table)  # the call reads the host arrays until it is collected
        self._check(self._lib.sst_explain_submit_f64(self._h, table._h, _p(m), _p(h), int(max_mods), P, _p(iv), _p(im), float(precision),
                                                     float(tolerance), 1 if with_memo else 0, _p(block), block.size))
        self._staged_P = P

    def explain_collect(self):
        """-> (status uint8[P], offsets uint32[P+1], records uint8[n, W]): views of this context's pinned buffers,
        valid until the next call on it."""
        m, h, iv, im, P, status, off, recs, table = self._submitted
        nc, W = C.c_uint64(), C.c_int()
        rc = self._lib.sst_explain_collect(self._h, table._h, C.byref(nc), C.byref(W))
        if rc == SST_ERR_NOMEM and nc.value and W.value:  # the pinned record buffer was too small: grow it, fetch from the device
            need = nc.value * W.value
            self._recs_hint = need + need // 4
            recs = self._pinned("recs", self._recs_hint)  # (the block is too small; the next submission allocates a larger one)
            self._last = (0, int(nc.value), int(W.value))
            off64 = np.empty(P + 1, dtype=np.uint64)
            self._check(self._lib.sst_explain_fetch(self._h, _p(status), _p(off64), _p(recs[: need])))
            off[:] = off64.astype(np.uint32)
        else:
            self._check(rc)
        self._submitted = None
        n, w = int(nc.value), int(W.value)
        self._last = (0, n, w)
        self._recs_hint = max(self.__dict__.get("_recs_hint", 0), n * w + n * w // 4)
        if rc == SST_OK:
            split, cap_n, hp = C.c_int(), C.c_uint64(), C.c_int()
            self._lib.sst_explain_rec_layout(self._h, C.byref(split), C.byref(cap_n), C.byref(hp))
            if split.value:  # planes: uint32 lo[cap_n], then hp byte planes of cap_n each
                cn = int(cap_n.value)
                lo = recs[: 4 * cn].view(np.uint32)[:n]
                return status, off, SplitRecords(lo, [recs[4 * cn + k * cn: 4 * cn + k * cn + n] for k in range(hp.value)])
        return status, off, recs[: n * w].reshape(n, w)

    # ---- N3 / N4 on a device-resident fragment frame (sst_ladder_*)
    def ladder_stage(self, su, observed, flags):
        s, o, f = _arr(su, np.float64), _arr(observed, np.float64), _arr(flags, np.uint8)
        if not (len(s) == len(o) == len(f)):
            raise ValueError("per-fragment arrays differ in length")
        self._check(self._lib.sst_ladder_stage(self._h, _p(s), _p(o), _p(f), len(s)))
        self._ladder_F = len(s)

    def ladder_round(self, table: "DeviceTable", max_weight: float, precision: float, tolerance: float, max_mods: int, ind, is_mod,
                     with_memo: bool = True):
        """-> (row mask as a Python int, number of calls, number of compositions) of one round over the alive fragments."""
        iv, im = _arr(ind, np.int32), _arr(is_mod, np.uint8)
        if len(iv) != table.R or len(im) != table.R:
            raise ValueError("ind / is_mod need one entry per table row")
        mask = np.zeros(4, dtype=np.uint32)
        nc, nk = C.c_uint64(), C.c_uint64()
        self._check(self._lib.sst_ladder_round(self._h, table._h, float(max_weight), float(precision), float(tolerance), int(max_mods), _p(iv), _p(im),
                                               1 if with_memo else 0, _p(mask), C.byref(nc), C.byref(nk)))
        self._staged_P = int(nc.value)
        self._last = (0, int(nk.value), int(self._lib.sst_explain_rec_width(self._h)))
        self._ladder_calls = int(nc.value)
        return sum(int(w) << (32 * k) for k, w in enumerate(mask)), int(nc.value), int(nk.value)

    def ladder_revalidate(self, table: "DeviceTable", precision: float, tolerance: float) -> int:
        n = C.c_int64()
        self._check(self._lib.sst_ladder_revalidate(self._h, table._h, float(precision), float(tolerance), C.byref(n)))
        return int(n.value)

    def ladder_fetch(self, calls: bool = True):
        """-> (alive uint8[F], keys float64[n], thresholds float64[n], call flags uint8[n]) of the last round."""
        alive = np.zeros(self._ladder_F, dtype=np.uint8)
        n = self._ladder_calls if calls else 0
        keys, thr, fl = np.zeros(n, dtype=np.float64), np.zeros(n, dtype=np.float64), np.zeros(n, dtype=np.uint8)
        self._check(self._lib.sst_ladder_fetch(self._h, _p(alive), _p(keys) if n else None, _p(thr) if n else None, _p(fl) if n else None))
        return alive, keys, thr, fl

    def explain_d2h_bytes(self) -> int:
        """Bytes the last collected submission copied device -> host."""
        return int(self._lib.sst_explain_d2h_bytes(self._h))

    def trace_ms(self, enable: bool) -> np.ndarray:
        """Device timeline (ms) of the last batch submitted on this context (sst_trace_ms)."""
        out = np.zeros(8, dtype=np.float32)
        self._check(self._lib.sst_trace_ms(self._h, 1 if enable else 0, _p(out)))
        return out

    def host_profile(self, enable: bool):
        """(ns[32], visits[32]) per section of the asynchronous entries since the last call (sst_host_profile)."""
        ns, calls = np.zeros(32, dtype=np.uint64), np.zeros(32, dtype=np.uint64)
        self._check(self._lib.sst_host_profile(1 if enable else 0, _p(ns), _p(calls)))
        return ns, calls

    def explain_phase_ns(self) -> np.ndarray:
        """Device timestamps of the last enumeration pass (see sst_explain_phase_ns)."""
        out = np.zeros(32, dtype=np.uint64)
        self._check(self._lib.sst_explain_phase_ns(self._h, _p(out)))
        return out

    def _pinned(self, name: str, nbytes: int) -> np.ndarray:
        """Grow-only page-locked staging buffer (uint8) kept on the context."""
        bufs = self.__dict__.setdefault("_pinned_bufs", {})
        buf = bufs.get(name)
        if buf is None or buf.size < nbytes:
            bufs[name] = buf = self.pinned_empty(max(nbytes + nbytes // 4, 4096), np.uint8)
        return buf

    def explain_fetch(self, want_records: bool = True, copy: bool = True):
        """-> (status uint8[P], offsets int64[P+1], records uint8[n, W] or None).

        Results land in pinned buffers owned by the context; ``copy=False`` hands out views of them
        (valid until the next fetch on this context)."""
        P = self._staged_P
        _nr, nc, W = self._last
        status = self._pinned("status", P)[:P]
        off = self._pinned("off", 8 * (P + 1))[: 8 * (P + 1)].view(np.uint64)
        recs = None
        if want_records:
            recs = self._pinned("recs", nc * W)[: nc * W].reshape(nc, W)
        self._check(self._lib.sst_explain_fetch(self._h, _p(status), _p(off), _p(recs) if (recs is not None and nc) else None))
        off = off.view(np.int64)
        if copy:
            return status.copy(), off.copy(), None if recs is None else recs.copy()
        return status, off, recs


class SplitRecords:
    """Records of a queued batch as they cross the bus (sst_set_record_split): ``lo`` uint32[n] holds the first four
    nucleotides of every composition (byte 0 = smallest row), ``planes[k]`` uint8[n] nucleotide 5 + k.  ``materialize()``
    gives the uint8[n, 8] array of whole records."""

    def __init__(self, lo: np.ndarray, planes):
        self.lo, self.planes = lo, list(planes)

    def __len__(self):
        return len(self.lo)

    def materialize(self) -> np.ndarray:
        n = len(self.lo)
        out = np.zeros((n, 8), dtype=np.uint8)
        if n:
            out[:, :4] = self.lo.view(np.uint8).reshape(n, 4)
            for k, p in enumerate(self.planes):
                out[:, 4 + k] = p
        return out


_BLOCK_LAYOUTS: Dict[int, tuple] = {}  # sst_explain_block_layout(P), remembered per batch size


class MemoFull(RuntimeError):
    pass


class DeviceTable:
    """Device-resident 2-bit table (+ mass-major row masks)."""

    def __init__(self, ctx: Context, handle, weights: np.ndarray):
        self.ctx, self._h, self.weights = ctx, handle, weights
        R, Cc, b, t = C.c_int(), C.c_int64(), C.c_float(), C.c_float()
        ctx._lib.sst_table_info(handle, C.byref(R), C.byref(Cc), C.byref(b), C.byref(t))
        self.R, self.C = R.value, Cc.value
        self._finalizer = weakref.finalize(self, ctx._lib.sst_table_destroy, ctx._h, handle)

    @property
    def limit(self) -> int:
        return self.C * 32

    def timings(self) -> Tuple[float, float]:
        b, t = C.c_float(), C.c_float()
        self.ctx._lib.sst_table_info(self._h, None, None, C.byref(b), C.byref(t))
        return float(b.value), float(t.value)

    def rebuild(self):
        self.ctx._check(self.ctx._lib.sst_table_rebuild(self.ctx._h, self._h))

    def download(self, out: Optional[np.ndarray] = None) -> np.ndarray:
        if out is None:
            out = np.empty((self.R, self.C), dtype=np.uint64)
        self.ctx._check(self.ctx._lib.sst_table_download(self.ctx._h, self._h, _p(out)))
        return out

    def download_masks(self, first_mass: int, n: int) -> np.ndarray:
        out = np.empty((n, 4), dtype=np.uint32)
        self.ctx._check(self.ctx._lib.sst_table_download_masks(self.ctx._h, self._h, int(first_mass), int(n), _p(out)))
        return out


_contexts: Dict[Tuple[int, int], Context] = {}


def default_device() -> int:
    return int(os.environ.get("SST_DEVICE", os.environ.get("LOCAL_RANK", "0")))


def context(device: Optional[int] = None, slot: int = 0) -> Context:
    """The context of a device.  ``slot`` > 0 gives further contexts on the same device (own streams, own scratch and
    result buffers): batches submitted on different slots are in flight together — the copies of one overlap the
    kernels of the other.  Tables built on slot 0 are used by every slot of the device."""
    d = default_device() if device is None else int(device)
    ctx = _contexts.get((d, int(slot)))
    if ctx is None:
        ctx = _contexts[(d, int(slot))] = Context(d)
    return ctx
