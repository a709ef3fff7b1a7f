"""The oracle's two line-search formulations against each other and against numpy.

LITERAL (default, what the GPU is compared with): Ceres' polynomial.cc as written — Vandermonde
system in the raw step sizes, full-pivot LU with Eigen's rank rule, critical points = real parts of
all eigenvalues of the balanced companion matrix (oracle/ref_shims/mini_ceres.cc).
FAST: normalised abscissae, divided differences, bracketed real roots polished by Newton — the
formulation csrc/lfr_math.cuh mirrors operation for operation.

The two must pick the same step wherever the Vandermonde system is well conditioned, and whole
solves must agree in both modes (positions to 1e-4 px, equal LM iteration counts) — including the
scenes whose line searches contract down to the minimum step size, where Ceres' own fit is
rank-truncated (see well_conditioned()).
"""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import get_problem
from oracle_util import (LS_FAST, LS_LITERAL, harvest_line_search_states, line_search_mode, minimize_interpolating,
                         well_conditioned)

TOL_UNITS = 1e-4 / 16.0


def test_companion_matrix_roots_match_numpy(oracle):
    rng = np.random.default_rng(11)
    re = np.zeros(8); im = np.zeros(8)
    for trial in range(2000):
        deg = int(rng.integers(3, 6))
        if trial % 3 == 0:      # prescribed real roots, some close together
            roots = rng.uniform(-2, 2, deg)
            if trial % 6 == 0:
                roots[1] = roots[0] + 1e-3
            co = np.poly(roots) * rng.uniform(0.1, 10)
        else:
            co = rng.normal(size=deg + 1) * 10.0 ** rng.uniform(-3, 3, size=deg + 1)
        co = np.ascontiguousarray(co, np.float64)
        n = oracle.lib.lfr_ref_find_polynomial_roots(co.ctypes.data, deg + 1, re.ctypes.data, im.ctypes.data)
        assert n == deg
        got = np.sort_complex(re[:n] + 1j * im[:n])
        want = np.sort_complex(np.roots(co))
        # compare as multisets through the polynomial they generate (root ordering of close pairs is arbitrary)
        scale = np.abs(want).max() + 1.0
        cond = 1e-5 if (trial % 6 == 0) else 1e-8   # a 1e-3 cluster of two roots is only sqrt(eps)-determined
        assert np.abs(got - want).max() <= cond * scale, (trial, got, want)


def test_quadratic_and_linear_closed_forms(oracle):
    re = np.zeros(4); im = np.zeros(4)
    co = np.array([2.0, -4.0])              # 2x - 4
    assert oracle.lib.lfr_ref_find_polynomial_roots(co.ctypes.data, 2, re.ctypes.data, im.ctypes.data) == 1 and re[0] == 2.0
    co = np.array([1.0, -3.0, 2.0])         # (x-1)(x-2)
    assert oracle.lib.lfr_ref_find_polynomial_roots(co.ctypes.data, 3, re.ctypes.data, im.ctypes.data) == 2
    assert sorted(re[:2].tolist()) == [1.0, 2.0] and not im[:2].any()
    co = np.array([1.0, 2.0, 5.0])          # -1 +- 2i
    assert oracle.lib.lfr_ref_find_polynomial_roots(co.ctypes.data, 3, re.ctypes.data, im.ctypes.data) == 2
    assert re[0] == -1.0 and re[1] == -1.0 and sorted(im[:2].tolist()) == [-2.0, 2.0]
    co = np.array([0.0, 0.0, 1.0, -3.0, 2.0])   # leading zeros are stripped
    assert oracle.lib.lfr_ref_find_polynomial_roots(co.ctypes.data, 5, re.ctypes.data, im.ctypes.data) == 2


def _random_states(rng, n, smallest):
    rows = []
    for k in range(n):
        three = k % 2 == 1
        x2 = 10.0 ** rng.uniform(np.log10(smallest / 0.02) if three else np.log10(smallest), 0)
        x2 = min(x2, 1.0)
        x1 = x2 * (rng.uniform(0.02, 0.6) if three else 1.0)
        f0, g0 = rng.normal(), -abs(rng.normal())
        f1, g1, f2, g2 = rng.normal(size=4) * np.array([1.0, 3.0 / x1, 1.0, 3.0 / x2])
        rows.append([f0, g0, x1, f1, g1, 1.0 if three else 0.0, x2 if three else 0.0, f2 if three else 0.0,
                     g2 if three else 0.0, 1e-3 * x1, 0.6 * x1])
    return np.ascontiguousarray(rows)


def _interpolant(row):
    f0, g0, x1, f1, g1, three, x2, f2, g2, lo, hi = row
    pts = [(0.0, f0, g0), (x1, f1, g1)] + ([(x2, f2, g2)] if three else [])
    h = max(p[0] for p in pts)
    deg = 2 * len(pts) - 1
    A, b = [], []
    for (x, f, g) in pts:
        t = x / h
        A.append([t ** k for k in range(deg, -1, -1)])
        b.append(f)
        A.append([k * t ** (k - 1) if k > 0 else 0.0 for k in range(deg, -1, -1)])
        b.append(g * h)
    return np.poly1d(np.linalg.solve(np.array(A), np.array(b))), h


def _same_choice(states, xa, xb, rtol):
    same = np.isclose(xa, xb, rtol=rtol, atol=0.0)
    for k in np.nonzero(~same)[0]:   # different abscissa only where the interpolant ties to rounding
        pl, h = _interpolant(states[k])
        scale = max(1.0, np.abs(pl.coeffs).max())
        assert abs(pl(xa[k] / h) - pl(xb[k] / h)) <= 1e-9 * scale, (k, xa[k], xb[k], states[k])
    return same.mean()


def test_literal_and_fast_pick_the_same_step_on_random_well_conditioned_states(oracle):
    states = _random_states(np.random.default_rng(99), 3000, smallest=0.02)
    assert well_conditioned(states).all()
    lit = minimize_interpolating(oracle, states, fast=False)
    fast = minimize_interpolating(oracle, states, fast=True)
    assert _same_choice(states, lit, fast, rtol=1e-9) >= 0.995


def test_literal_fit_is_rank_truncated_for_tiny_steps(oracle):
    """Documents WHY the comparison is restricted: with samples at ~1e-3 the raw-step Vandermonde
    system loses numerical rank under Eigen's rule and Ceres' own polynomial is no longer the
    interpolant — the literal mode reproduces that, the fast formulation does not."""
    row = np.array([[0.7, -0.5, 5e-4, 0.6999, -0.3, 1.0, 1.5e-3, 0.6995, 0.1, 5e-7, 3e-4]])
    lit = minimize_interpolating(oracle, row, fast=False)[0]
    fast = minimize_interpolating(oracle, row, fast=True)[0]
    pl, h = _interpolant(row[0])
    grid = np.linspace(row[0, 9], row[0, 10], 20001)
    assert abs(pl(fast / h) - pl(grid / h).min()) <= 1e-9      # fast = the true interpolant's minimiser
    assert row[0, 9] <= lit <= row[0, 10]


SCENES = [("cfg1", 1.0), ("cfg2", 1.0), ("cfg3", 0.5), ("cfg4", 0.5), ("ring60", 1.0)]


@pytest.mark.parametrize("name,scale", SCENES)
def test_solves_agree_in_both_modes_and_steps_agree_on_harvested_states(oracle, name, scale):
    _, p = get_problem(name, scale=scale)
    opt = oracle.default_options(n_threads=8)
    with line_search_mode(oracle, LS_LITERAL):
        pos_l, st_l, states = harvest_line_search_states(oracle, p, opt)
    with line_search_mode(oracle, LS_FAST):
        pos_f, st_f = oracle.solve(p, opt)
    assert np.abs(pos_l - pos_f).max() <= TOL_UNITS
    assert np.array_equal(st_l["iterations"], st_f["iterations"])
    assert np.array_equal(st_l["termination"], st_f["termination"])
    if states.shape[0]:
        wc = well_conditioned(states)
        lit = minimize_interpolating(oracle, states[wc], fast=False)
        fast = minimize_interpolating(oracle, states[wc], fast=True)
        if wc.any():
            assert _same_choice(states[wc], lit, fast, rtol=1e-9) >= 0.99


def test_linesearch_fixture_agrees_in_both_modes(oracle):
    from lfr_b200 import build_problem, wire
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "linesearch_matches.pb")
    with open(path, "rb") as fh:
        p = build_problem(wire.decode_matching_file(fh.read()))
    with line_search_mode(oracle, LS_LITERAL):
        pos_l, st_l, states = harvest_line_search_states(oracle, p)
    with line_search_mode(oracle, LS_FAST):
        pos_f, st_f = oracle.solve(p, oracle.default_options(n_threads=1))
    assert states.shape[0] > 20
    assert np.abs(pos_l - pos_f).max() <= TOL_UNITS
    assert np.array_equal(st_l["iterations"], st_f["iterations"])
