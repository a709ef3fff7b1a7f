"""Mismatches between the grid root finder, the derivative recursion (GPU) and the oracle."""
import ctypes as C
import os
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np  # noqa: E402

from lfr_b200.capi import load_b200  # noqa: E402
from oracle_util import load_oracle  # noqa: E402
import test_gpu_parity as T  # noqa: E402

lib, oracle = load_b200(), load_oracle()
n = 6000
coef, lohi = T._quartic_cases(np.random.default_rng(99), n)
f = lib.lib.lfr_debug_quartic_roots
f.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p]
res = {}
for g in (0, 1):
    roots = np.zeros((n, 4)); cnt = np.zeros(n, dtype=np.int32)
    assert f(coef.ctypes.data, lohi.ctypes.data, n, g, roots.ctypes.data, cnt.ctypes.data) == 0
    res[g] = (roots, cnt)
shown = 0
nm = [0, 0]
for k in range(n):
    out = np.zeros(8)
    c = oracle.lib.lfr_ref_polynomial_roots(coef[k].ctypes.data, 5, float(lohi[k, 0]), float(lohi[k, 1]), out.ctypes.data)
    o = out[:c]
    for g in (0, 1):
        r = res[g][0][k][:res[g][1][k]]
        ok = len(r) == c and np.allclose(r, o, rtol=1e-9, atol=1e-12)
        if not ok:
            nm[g] += 1
            if shown < 12:
                shown += 1
                npr = np.roots(coef[k])
                print("case", k, "mode", k % 6, "grid" if g else "recursion", "interval", lohi[k])
                print("   gpu   ", r)
                print("   oracle", o)
                print("   numpy ", np.sort_complex(npr))
print("mismatches: recursion", nm[0], "grid", nm[1], "of", n)
