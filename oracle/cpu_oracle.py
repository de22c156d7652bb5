"""ctypes front-end of ``oracle/knpemi_oracle.c`` (TEST INFRASTRUCTURE).

``rhs``  : one right-hand-side evaluation of a restated model.
``step`` : one PDE step of scheme O1 (RK4 x n_sub + current epilogue) over AoS
           tables, i.e. the row loop of reference src/knpemi/odeSolver.py:106-123
           with the fixed-step integrator of SURVEY.md 8(c).
``step_fn``: the same driver over an arbitrary ``void(double,double*,double*,double*)``
           function pointer -- used by tests/golden/make_golden.py to push the
           *reference's own* numba cfuncs through the identical scheme.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libknpemi_oracle.so")
_lib = None

MODEL_NAMES = ("hh_ideal", "hh_tissue", "glial_tissue", "glial_bench",
               "calibration", "hh_test")

_DP = ctypes.POINTER(ctypes.c_double)


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (oracle/Makefile). Returns the .so path."""
    src = os.path.join(_HERE, "knpemi_oracle.c")
    if (force or not os.path.exists(_LIB_PATH)
            or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)):
        subprocess.run(["make", "-C", _HERE, "-s", "-B"], check=True)
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        L.kemo_model_find.argtypes = [ctypes.c_char_p]
        L.kemo_model_find.restype = ctypes.c_int
        L.kemo_model_dims.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_int),
                                      ctypes.POINTER(ctypes.c_int)]
        L.kemo_rhs.argtypes = [ctypes.c_int, ctypes.c_double, _DP, _DP, _DP]
        L.kemo_step.argtypes = [ctypes.c_int, ctypes.c_int64, _DP, _DP,
                                ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_int]
        L.kemo_step.restype = ctypes.c_int64
        L.kemo_step_fn.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int64,
                                   _DP, _DP, ctypes.c_double, ctypes.c_double,
                                   ctypes.c_int, ctypes.c_int]
        L.kemo_step_fn.restype = ctypes.c_int64
        L.kemo_step_dp45.argtypes = [ctypes.c_int, ctypes.c_int64, _DP, _DP, _DP, ctypes.c_double,
                                     ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_int,
                                     ctypes.POINTER(ctypes.c_int64)]
        L.kemo_step_dp45.restype = ctypes.c_int64
        L.kemo_max_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def model_id(name: str) -> int:
    mid = lib().kemo_model_find(name.encode())
    if mid < 0:
        raise KeyError(f"oracle has no model {name!r}")
    return mid


def dims(name: str) -> tuple[int, int]:
    ns, np_ = ctypes.c_int(), ctypes.c_int()
    lib().kemo_model_dims(model_id(name), ctypes.byref(ns), ctypes.byref(np_))
    return ns.value, np_.value


def _ptr(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(_DP)


def rhs(name: str, t: float, y: np.ndarray, p: np.ndarray):
    """Return (dy, p_after) for one evaluation; inputs are not modified."""
    ns, np_ = dims(name)
    y = np.ascontiguousarray(y, dtype=np.float64).copy()
    p = np.ascontiguousarray(p, dtype=np.float64).copy()
    assert y.shape == (ns,) and p.shape == (np_,)
    dy = np.zeros(ns)
    lib().kemo_rhs(model_id(name), float(t), _ptr(y), _ptr(dy), _ptr(p))
    return dy, p


def step(name: str, states: np.ndarray, params: np.ndarray, t0: float, dt: float,
         n_sub: int = 25, n_threads: int = 0) -> int:
    """Advance AoS tables IN PLACE by one PDE step; returns #non-finite rows."""
    ns, np_ = dims(name)
    assert states.ndim == 2 and states.shape[1] == ns
    assert params.shape == (states.shape[0], np_)
    return int(lib().kemo_step(model_id(name), states.shape[0], _ptr(states), _ptr(params),
                               float(t0), float(dt), int(n_sub), int(n_threads)))


def step_fn(address: int, states: np.ndarray, params: np.ndarray, t0: float, dt: float,
            n_sub: int = 25, n_threads: int = 1) -> int:
    """Scheme O1 over a foreign RHS function pointer (e.g. a numba cfunc)."""
    return int(lib().kemo_step_fn(ctypes.c_void_p(address), states.shape[1], params.shape[1],
                                  states.shape[0], _ptr(states), _ptr(params),
                                  float(t0), float(dt), int(n_sub), int(n_threads)))


def step_dp45(name: str, states: np.ndarray, params: np.ndarray, hsug: np.ndarray, t0: float, dt: float,
              rtol: float = 1e-8, atol: float = 1e-10, n_threads: int = 0):
    """Scheme O3 (Dormand-Prince 5(4), error-controlled) IN PLACE; returns (#failed rows,
    accepted steps, rejected steps).  `hsug` is the per-row warm-start step size (in/out)."""
    ns, np_ = dims(name)
    assert states.shape[1] == ns and params.shape == (states.shape[0], np_) and hsug.shape == (states.shape[0],)
    stats = (ctypes.c_int64 * 2)()
    bad = lib().kemo_step_dp45(model_id(name), states.shape[0], _ptr(states), _ptr(params), _ptr(hsug),
                               float(t0), float(dt), float(rtol), float(atol), int(n_threads), stats)
    return int(bad), int(stats[0]), int(stats[1])


def max_threads() -> int:
    return int(lib().kemo_max_threads())
