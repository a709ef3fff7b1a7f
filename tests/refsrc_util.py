"""Test-side access to oracle/_ref/libref_cost.so — the REFERENCE's own cost.cc compiled against
the shim headers (oracle/build_ref.py).  Test infrastructure only."""
import ctypes as C
import importlib.util
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_lib = None


def _build_mod():
    spec = importlib.util.spec_from_file_location("lfr_build_ref", os.path.join(ROOT, "oracle", "build_ref.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_refsrc():
    """Returns the ctypes library, or None when it is neither built nor buildable here."""
    global _lib
    if _lib is None:
        try:
            path = _build_mod().build_cost()
        except Exception:
            return None
        L = C.CDLL(path)
        for name, n in (("lfr_refsrc_interpolate", 6), ("lfr_refsrc_cost", 7), ("lfr_refsrc_residual", 5)):
            getattr(L, name).argtypes = [C.c_uint64] + [C.c_void_p] * (n - 1)
            getattr(L, name).restype = None
        _lib = L
    return _lib


def ref_interpolate(L, grids, rc):
    n = grids.shape[0]
    grids = np.ascontiguousarray(grids, np.float64)
    rc = np.ascontiguousarray(rc, np.float64)
    f = np.zeros((n, 2)); dr = np.zeros((n, 2)); dc = np.zeros((n, 2))
    L.lfr_refsrc_interpolate(n, grids.ctypes.data, rc.ctypes.data, f.ctypes.data, dr.ctypes.data, dc.ctypes.data)
    return f, dr, dc


def ref_cost(L, grids, x1, x2):
    n = grids.shape[0]
    grids = np.ascontiguousarray(grids, np.float64)
    x1 = np.ascontiguousarray(x1, np.float64)
    x2 = np.ascontiguousarray(x2, np.float64)
    r = np.zeros((n, 2)); j1 = np.zeros((n, 4)); j2 = np.zeros((n, 4))
    L.lfr_refsrc_cost(n, grids.ctypes.data, x1.ctypes.data, x2.ctypes.data, r.ctypes.data, j1.ctypes.data, j2.ctypes.data)
    return r, j1, j2


def cost_cases(n, seed=0):
    """Random, clamped and boundary inputs (grids are fp32 values widened, as solve.cc:460-472 does)."""
    rng = np.random.default_rng(seed)
    grids = rng.uniform(-0.6, 0.6, (n, 18)).astype(np.float32)
    x1 = rng.uniform(-1.0, 1.0, (n, 2))
    k = max(n // 20, 1)
    x1[0 * k:1 * k, 0] = 0.5
    x1[1 * k:2 * k, 1] = -0.5
    x1[2 * k:3 * k] = 0.0
    x1[3 * k:4 * k, 0] = np.nextafter(0.5, 1.0)
    x1[4 * k:5 * k] = [0.5, -0.5]
    x1[5 * k:6 * k, 1] = np.nextafter(-0.5, -1.0)
    x1[6 * k:7 * k] = rng.uniform(-0.5, 0.5, (k, 2))
    x2 = rng.uniform(-1.0, 1.0, (n, 2))
    return grids, x1, x2
