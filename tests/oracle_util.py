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
        _oracle = lib
    return _oracle
