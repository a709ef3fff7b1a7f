"""CPU model of the CUDA line search's Budan-Fourier grid isolation
(csrc/lfr_math.cuh :: quartic_roots_grid): the per-cell decision it takes from
sign-variation counts is checked against the oracle's roots
(lfr_ref_polynomial_roots, derivative recursion), so the argument the kernel
relies on is exercised without a GPU.  The kernel itself is compared with the
oracle in tests/test_gpu_parity.py::test_quartic_root_finders_agree."""
import numpy as np

from test_gpu_parity import _quartic_cases


def taylor(q, t):
    """(q, q', q''/2, q'''/6, q''''/24) at t by repeated synthetic division."""
    a = list(q)
    for last in (4, 3, 2, 1):
        for i in range(1, last + 1):
            a[i] = a[i] + a[i - 1] * t
    return a[4], a[3], a[2], a[1], a[0]


def classify(q, lo, hi):
    """Per cell: 0 none, 1 one root, 2 one extremum to examine, 3 undecided; None -> fallback."""
    w = (hi - lo) / 31.0
    ts = [hi if L == 31 else lo + w * L for L in range(32)]
    V, Vp = [], []
    for t in ts:
        s = taylor(q, t)
        if not np.isfinite(s[0]) or any(v == 0.0 for v in s[:4]):
            return None, ts
        n = [bool(v < 0) for v in s]
        vp = int(n[1] != n[2]) + int(n[2] != n[3]) + int(n[3] != n[4])
        Vp.append(vp)
        V.append(int(n[0] != n[1]) + vp)
    kinds = []
    for L in range(31):
        D, Dp = V[L] - V[L + 1], Vp[L] - Vp[L + 1]
        if D < 0 or Dp < 0 or D > 3:
            k = 3
        elif D == 0:
            k = 0
        elif D == 1:
            k = 1
        elif D == 2:
            k = 0 if Dp == 0 else (2 if Dp == 1 else 3)
        else:
            k = 1 if Dp == 0 else 3
        kinds.append(k)
    if 3 in kinds:
        return None, ts
    return kinds, ts


def test_cell_decisions_match_the_oracles_roots(oracle):
    rng = np.random.default_rng(99)
    coef, lohi = _quartic_cases(rng, 3000)
    decided = cells_one = cells_two = 0
    for k in range(coef.shape[0]):
        q, (lo, hi) = coef[k], lohi[k]
        kinds, ts = classify(q, float(lo), float(hi))
        if kinds is None:
            continue
        decided += 1
        out = np.zeros(8)
        n = oracle.lib.lfr_ref_polynomial_roots(q.ctypes.data, 5, float(lo), float(hi), out.ctypes.data)
        roots = out[:n]
        for L, kind in enumerate(kinds):
            inside = int(((roots > ts[L]) & (roots <= ts[L + 1])).sum())
            if kind == 0:
                # a pair closer than the oracle can separate may hide here; a lone root may not
                assert inside in (0, 2), (k, L, inside)
                if inside == 2:
                    r = roots[(roots > ts[L]) & (roots <= ts[L + 1])]
                    assert abs(r[1] - r[0]) <= 1e-6 * max(1.0, abs(r[0]))
            elif kind == 1:
                assert inside == 1, (k, L, inside)
                cells_one += 1
            else:
                assert inside in (0, 2), (k, L, inside)
                cells_two += inside == 2
    # the grid decides most cases on its own and the two-roots-in-a-cell branch is exercised
    assert decided > 0.8 * coef.shape[0] and cells_one > 1000 and cells_two > 50, (decided, cells_one, cells_two)
