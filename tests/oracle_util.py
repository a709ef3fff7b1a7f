"""Test-side access to the CPU oracle (oracle/liblfr_ref.so).  Only tests/,
__graft_entry__.smoke() and bench.py's CPU-baseline legs may use this."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

_oracle = None


def load_oracle():
    global _oracle
    if _oracle is None:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import importlib.util
        spec = importlib.util.spec_from_file_location("lfr_oracle_build", os.path.join(ROOT, "oracle", "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        path = mod.build()
        from lfr_b200.capi import Library
        lib = Library(path)
        assert lib.backend == "cpu-oracle"
        L = lib.lib
        L.lfr_ref_interpolate.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
        L.lfr_ref_interpolate.restype = None
        L.lfr_ref_loss.argtypes = [C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_void_p]
        L.lfr_ref_loss.restype = None
        L.lfr_ref_minimize_interpolating_polynomial.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double,
                                                                 C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.lfr_ref_minimize_interpolating_polynomial.restype = None
        L.lfr_ref_polynomial_roots.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_void_p]
        L.lfr_ref_polynomial_roots.restype = C.c_int
        L.lfr_ref_minimize_interpolating_polynomial_fast.argtypes = L.lfr_ref_minimize_interpolating_polynomial.argtypes
        L.lfr_ref_minimize_interpolating_polynomial_fast.restype = None
        L.lfr_ref_set_line_search_mode.argtypes = [C.c_int]
        L.lfr_ref_set_line_search_mode.restype = None
        L.lfr_ref_get_line_search_mode.restype = C.c_int
        L.lfr_ref_find_polynomial_roots.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.lfr_ref_find_polynomial_roots.restype = C.c_int
        L.lfr_ref_harvest.argtypes = [C.c_int]
        L.lfr_ref_harvest.restype = None
        L.lfr_ref_harvest_take.argtypes = [C.c_void_p, C.c_uint64]
        L.lfr_ref_harvest_take.restype = C.c_uint64
        _oracle = lib
    return _oracle


LS_LITERAL, LS_FAST = 0, 1


class line_search_mode:
    """with line_search_mode(orc, LS_FAST): ... — the oracle's interpolation mode (default: literal)."""

    def __init__(self, orc, mode):
        self.orc, self.mode = orc, mode

    def __enter__(self):
        self.prev = self.orc.lib.lfr_ref_get_line_search_mode()
        self.orc.lib.lfr_ref_set_line_search_mode(self.mode)
        return self

    def __exit__(self, *exc):
        self.orc.lib.lfr_ref_set_line_search_mode(self.prev)
        return False


def harvest_line_search_states(orc, problem, options=None):
    """Solve `problem` with the oracle and return (positions, stats, states[n, 11]): every
    interpolation state the line search went through."""
    import numpy as np
    orc.lib.lfr_ref_harvest(1)
    try:
        pos, st = orc.solve(problem, options if options is not None else orc.default_options(n_threads=1))
        n = int(orc.lib.lfr_ref_harvest_take(None, 0))
        out = np.zeros((n, 11), dtype=np.float64)
        if n:
            orc.lib.lfr_ref_harvest_take(out.ctypes.data, n)
    finally:
        orc.lib.lfr_ref_harvest(0)
    return pos, st, out


def minimize_interpolating(orc, states, fast):
    """Step sizes chosen for line-search states [n, 11] by the literal (fast=False) or the fast
    formulation of the oracle."""
    import numpy as np
    fn = orc.lib.lfr_ref_minimize_interpolating_polynomial_fast if fast else orc.lib.lfr_ref_minimize_interpolating_polynomial
    out = np.zeros(states.shape[0])
    x, v = C.c_double(), C.c_double()
    with line_search_mode(orc, LS_FAST if fast else LS_LITERAL):
        for k, r in enumerate(states):
            three = r[5] != 0.0
            smp = [[0.0, r[0], r[1], 1, 1], [r[2], r[3], r[4], 1, 1]] + ([[r[6], r[7], r[8], 1, 1]] if three else [])
            a = np.array(smp, dtype=np.float64)
            fn(a.ctypes.data, len(smp), float(r[9]), float(r[10]), C.byref(x), C.byref(v))
            out[k] = x.value
    return out


def well_conditioned(states, min_step=0.02):
    """States whose raw-step Vandermonde system (polynomial.cc) keeps full numerical rank in Eigen's
    fullPivLu: the smallest sample abscissa is not tiny.  Below ~2e-3 the x^5 column falls under the
    rank threshold (eps * 6 * max pivot) and Ceres itself fits a rank-truncated polynomial."""
    import numpy as np
    three = states[:, 5] != 0.0
    smallest = np.where(three, np.minimum(states[:, 2], states[:, 6]), states[:, 2])
    return smallest >= min_step
