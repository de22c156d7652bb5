"""ctypes binding of libknpemi_b200.so (C ABI: include/knpemi_b200.h).

This is the thin layer the north star asks for: NumPy buffers are passed as raw
pointers (``arr.ctypes.data``); there is no PyTorch and no CPU fallback -- if the
shared library is missing or no CUDA device is present the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .codegen.build import RUNTIME_LIB

KEM_STATE, KEM_PARAM = 0, 1
KEM_SCHEME_RK4 = 0
KEM_SCHEME_DP45 = 1
KEM_NONFINITE = 1
KEM_STEP_FAILED = 2
KEM_MAX_STIM = 4
UNREAD_POLICIES = {"auto": 0, "shadow": 1, "upload": 2, "discard": 3}


class KemError(RuntimeError):
    """Argument / CUDA / model error reported by libknpemi_b200."""


class NonFiniteStateError(AssertionError):
    """A membrane state became non-finite during a step.

    Subclass of AssertionError: the reference signals integration failure with
    ``assert success`` (src/knpemi/odeSolver.py:121)."""


class StepControlError(AssertionError):
    """Scheme "dp45" could not reach t+dt within its step limit / minimum step size.

    Subclass of AssertionError for the same reason as :class:`NonFiniteStateError`."""


class kem_model_info(C.Structure):
    _fields_ = [("ns", C.c_int), ("np", C.c_int), ("n_out", C.c_int), ("n_used", C.c_int),
                ("n_tslots", C.c_int), ("out_cols", C.c_int * 64), ("name", C.c_char * 64),
                ("source_hash", C.c_char * 32)]


class kem_step_times(C.Structure):
    _fields_ = [("ms_h2d", C.c_double), ("ms_kernel", C.c_double), ("ms_d2h", C.c_double),
                ("ms_total", C.c_double)]


class kem_io_column(C.Structure):
    _fields_ = [("kind", C.c_int), ("col", C.c_int), ("host", C.c_void_p)]


_DP = C.POINTER(C.c_double)
_IP = C.POINTER(C.c_int)
_H = C.c_void_p

# name -> (restype, argtypes); every symbol include/knpemi_b200.h declares
SIGNATURES = {
    "kem_version": (C.c_int, []),
    "kem_last_error": (C.c_char_p, []),
    "kem_device_count": (C.c_int, [_IP]),
    "kem_model_load": (C.c_int, [C.c_char_p, _IP]),
    "kem_model_find": (C.c_int, [C.c_char_p, _IP]),
    "kem_model_get_info": (C.c_int, [C.c_int, C.POINTER(kem_model_info)]),
    "kem_model_launch_info": (C.c_int, [C.c_int, C.c_int, C.c_int, _IP, _IP]),
    "kem_create": (C.c_int, [C.c_int, C.c_int64, C.c_int, _IP, _DP, _DP, C.POINTER(_H)]),
    "kem_destroy": (C.c_int, [_H]),
    "kem_n_dof": (C.c_int, [_H, C.POINTER(C.c_int64)]),
    "kem_shard_range": (C.c_int, [_H, C.c_int, _IP, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "kem_set_uniform": (C.c_int, [_H, C.c_int, C.c_int, C.c_double]),
    "kem_set_column": (C.c_int, [_H, C.c_int, C.c_int, C.c_void_p, C.c_int64]),
    "kem_set_column_masked": (C.c_int, [_H, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int64]),
    "kem_set_value_masked": (C.c_int, [_H, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_int64]),
    "kem_get_column": (C.c_int, [_H, C.c_int, C.c_int, C.c_void_p, C.c_int64]),
    "kem_column_is_uniform": (C.c_int, [_H, C.c_int, C.c_int, _IP, _DP]),
    "kem_column_location": (C.c_int, [_H, C.c_int, C.c_int, _IP]),
    "kem_set_stimulus_mask": (C.c_int, [_H, C.c_void_p, C.c_int64]),
    "kem_step": (C.c_int, [_H, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, _IP, _DP, _IP]),
    "kem_step_timed": (C.c_int, [_H, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, _IP, _DP, _IP,
                                 C.POINTER(kem_step_times)]),
    "kem_step_io": (C.c_int, [_H, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, _IP, _DP,
                              C.c_int, C.POINTER(kem_io_column), C.c_int, C.POINTER(kem_io_column),
                              _IP, C.POINTER(kem_step_times)]),
    "kem_sync": (C.c_int, [_H]),
    "kem_set_unread_policy": (C.c_int, [_H, C.c_int]),
    "kem_set_step_chunks": (C.c_int, [_H, C.c_int]),
    "kem_set_io_tuning": (C.c_int, [_H, C.c_int, C.c_int]),
    "kem_plan_chunks": (C.c_int, [C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                  C.c_int, _IP]),
    "kem_set_tolerances": (C.c_int, [_H, C.c_double, C.c_double]),
    "kem_set_activity_sort": (C.c_int, [_H, C.c_int]),
    "kem_get_step_stats": (C.c_int, [_H, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "kem_timer_begin": (C.c_int, [_H]),
    "kem_timer_end": (C.c_int, [_H, _DP]),
    "kem_set_block": (C.c_int, [_H, C.c_int]),
    "kem_launch_count": (C.c_int, [_H, C.POINTER(C.c_int64)]),
    "kem_device_map_set": (C.c_int, [_H, C.c_int, C.c_void_p, C.c_int64]),
    "kem_device_gather": (C.c_int, [_H, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]),
    "kem_device_scatter": (C.c_int, [_H, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]),
    "kem_device_gather_diff": (C.c_int, [_H, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                         C.c_int]),
    "kem_device_affine_combine": (C.c_int, [C.c_int, C.c_int64, C.c_void_p, C.c_double, C.c_int, _DP,
                                            C.POINTER(C.c_void_p)]),
    "kem_device_gather_affine": (C.c_int, [_H, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, _DP,
                                           C.POINTER(C.c_void_p), C.c_int]),
    "kem_device_copy_in": (C.c_int, [_H, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "kem_device_copy_out": (C.c_int, [_H, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "kem_device_alloc": (C.c_int, [C.c_int, C.c_size_t, C.POINTER(C.c_void_p)]),
    "kem_device_free": (C.c_int, [C.c_int, C.c_void_p]),
    "kem_device_upload": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]),
    "kem_device_download": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]),
    "kem_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t]),
    "kem_host_free": (C.c_int, [C.c_void_p]),
    "kem_host_register": (C.c_int, [C.c_void_p, C.c_size_t]),
    "kem_host_unregister": (C.c_int, [C.c_void_p]),
    "kem_host_is_pinned": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_int)]),
    "kem_fp64_peak": (C.c_int, [C.c_int, _DP, _DP]),
    "kem_hbm_copy_peak": (C.c_int, [C.c_int, _DP]),
    "kem_link_probe": (C.c_int, [C.c_int, C.c_size_t, C.c_size_t, C.c_int, C.c_int, _DP, _DP]),
}

_lib = None


def library_path() -> str:
    return os.environ.get("KNPEMI_B200_LIB", RUNTIME_LIB)


def lib() -> C.CDLL:
    """Load libknpemi_b200.so (built by ``__graft_entry__.build()``); raises if absent."""
    global _lib
    if _lib is None:
        path = library_path()
        if not os.path.exists(path):
            raise KemError(f"{path} not found: build it with `python -c 'import __graft_entry__ as g; "
                           "g.build()'` (there is no CPU fallback)")
        L = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(rc: int, what: str = "") -> int:
    if rc == 0:
        return 0
    msg = lib().kem_last_error().decode(errors="replace")
    if rc == KEM_NONFINITE:
        raise NonFiniteStateError(msg or "non-finite membrane state")
    if rc == KEM_STEP_FAILED:
        raise StepControlError(msg or "error-controlled step failed")
    raise KemError(f"{what or 'libknpemi_b200'} failed (code {rc}): {msg}")


_from_buffer, _addressof = C.c_char.from_buffer, C.addressof


def address(a: np.ndarray) -> int:
    """Data pointer of a NumPy array, three times faster than ``a.ctypes.data`` (0.6 instead of
    1.9 us: on a membrane of a few hundred DOFs the eleven setter / getter calls of a PDE step
    are mostly interpreter time)."""
    try:
        return _addressof(_from_buffer(a))
    except (TypeError, ValueError, BufferError):          # read-only, empty or strided array
        return a.ctypes.data


def device_count() -> int:
    n = C.c_int(0)
    rc = lib().kem_device_count(C.byref(n))
    return n.value if rc == 0 else 0


def load_model(so_path: str) -> int:
    mid = C.c_int(-1)
    check(lib().kem_model_load(so_path.encode(), C.byref(mid)), "kem_model_load")
    return mid.value


def model_info(model_id: int) -> kem_model_info:
    info = kem_model_info()
    check(lib().kem_model_get_info(model_id, C.byref(info)), "kem_model_get_info")
    return info


def _free_pinned(ptr: int):
    try:
        lib().kem_host_free(C.c_void_p(ptr))
    except Exception:
        pass


def pinned_empty(n: int, dtype=np.float64) -> np.ndarray:
    """1-D NumPy array over page-locked host memory (kem_host_alloc).

    The memory belongs to the array: it is released when the array (and every view of it)
    is garbage collected, so `pinned_empty(n)` is as safe to pass around as `np.empty(n)`."""
    import weakref
    dtype = np.dtype(dtype)
    nbytes = max(int(n) * dtype.itemsize, 8)
    p = C.c_void_p()
    check(lib().kem_host_alloc(C.byref(p), nbytes), "kem_host_alloc")
    buf = (C.c_char * nbytes).from_address(p.value)
    weakref.finalize(buf, _free_pinned, p.value)       # the array's base keeps `buf` alive
    return np.frombuffer(buf, dtype=dtype, count=int(n))


class PinnedArray:
    """Thin holder kept for callers that want an explicit object: `.array` is `pinned_empty(n)`."""

    def __init__(self, n: int, dtype=np.float64):
        self.array = pinned_empty(n, dtype)
        self.dtype = self.array.dtype
        self.nbytes = self.array.nbytes

    def free(self):
        self.array = None


def host_is_pinned(a) -> bool:
    """True if transfers from/to the ndarray `a` take the direct (page-locked) path."""
    out = C.c_int(0)
    check(lib().kem_host_is_pinned(C.c_void_p(a.ctypes.data), a.nbytes, C.byref(out)), "kem_host_is_pinned")
    return bool(out.value)


def fp64_peak(dev: int = 0) -> tuple[float, float]:
    tf, ms = C.c_double(), C.c_double()
    check(lib().kem_fp64_peak(dev, C.byref(tf), C.byref(ms)), "kem_fp64_peak")
    return tf.value, ms.value


def hbm_copy_peak(dev: int = 0) -> float:
    g = C.c_double()
    check(lib().kem_hbm_copy_peak(dev, C.byref(g)), "kem_hbm_copy_peak")
    return g.value


def link_probe(dev: int, nbytes: int, reps_h2d: int, reps_d2h: int, span: int = 0) -> tuple[float, float]:
    """(GB/s host->device, GB/s device->host) of `reps_*` concurrent pinned copies of `nbytes`
    walking through `span` bytes of host memory per direction (0: one cache-resident buffer)."""
    a, b = C.c_double(), C.c_double()
    check(lib().kem_link_probe(dev, nbytes, span, reps_h2d, reps_d2h, C.byref(a), C.byref(b)), "kem_link_probe")
    return (nbytes * reps_h2d / (a.value * 1e-3) / 1e9 if reps_h2d else 0.0,
            nbytes * reps_d2h / (b.value * 1e-3) / 1e9 if reps_d2h else 0.0)


def link_ceiling(dev: int = 0, nbytes: int = 5 << 20, reps: int = 96, span: int = 480 << 20) -> dict:
    """Measured pinned-copy ceilings of one GPU's host link, GB/s per direction: each direction
    alone, both saturated, and each direction while the other one runs for twice as long.
    Defaults: 5 MB copies (one column chunk of the pipelined exchange at 1e7 DOFs) streaming
    through 480 MB of host memory per direction (far above the last-level cache)."""
    h_alone, _ = link_probe(dev, nbytes, reps, 0, span)
    _, d_alone = link_probe(dev, nbytes, 0, reps, span)
    h_both, d_both = link_probe(dev, nbytes, reps, reps, span)
    h_loaded, _ = link_probe(dev, nbytes, reps, 2 * reps, span)
    _, d_loaded = link_probe(dev, nbytes, 2 * reps, reps, span)
    return {"h2d_alone": h_alone, "d2h_alone": d_alone, "h2d_both": h_both, "d2h_both": d_both,
            "h2d_under_d2h": h_loaded, "d2h_under_h2d": d_loaded, "copy_bytes": nbytes, "reps": reps,
            "span_bytes": span}


class DeviceArray:
    """float64 buffer in the HBM of one device (kem_device_alloc) -- stands in for a PDE
    coefficient vector that already lives on the GPU (rows f1/f3 of SURVEY.md 8f)."""

    def __init__(self, dev: int, host: np.ndarray):
        host = np.ascontiguousarray(host, dtype=np.float64)
        self.dev, self.n = dev, len(host)
        p = C.c_void_p()
        check(lib().kem_device_alloc(dev, host.nbytes, C.byref(p)), "kem_device_alloc")
        self.ptr = p.value
        check(lib().kem_device_upload(dev, C.c_void_p(self.ptr), host.ctypes.data, host.nbytes),
              "kem_device_upload")

    def to_host(self) -> np.ndarray:
        out = np.empty(self.n, dtype=np.float64)
        check(lib().kem_device_download(self.dev, out.ctypes.data, C.c_void_p(self.ptr), out.nbytes),
              "kem_device_download")
        return out

    def free(self):
        if self.ptr:
            lib().kem_device_free(self.dev, C.c_void_p(self.ptr))
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
