"""Committed fixtures (tests/golden/, made by tests/golden/make_golden.py from
the oracle — the reference has no golden vectors of its own)."""
import os

import numpy as np
import pytest

from lfr_b200 import build_problem, wire
from lfr_b200.solver import assemble_solution

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["tiny", "linesearch", "fountain_2pct"]
TOL_UNITS = 1e-4 / 16.0


def load(name):
    with open(os.path.join(GOLD, "%s_matches.pb" % name), "rb") as fh:
        ms = wire.decode_matching_file(fh.read())
    return build_problem(ms), np.load(os.path.join(GOLD, "%s_expected.npz" % name))


@pytest.mark.parametrize("name", CASES)
def test_host_stage_and_oracle_reproduce_golden(oracle, name):
    p, exp = load(name)
    for k in ("track", "comp", "is_root", "comp_ptr", "comp_nodes"):
        assert np.array_equal(getattr(p, k), exp[k]), k
    pos, st = oracle.solve(p, oracle.default_options(n_threads=3))
    np.testing.assert_allclose(pos, exp["positions"], rtol=0, atol=1e-12)
    assert np.array_equal(st["iterations"], exp["iterations"])
    assert np.array_equal(st["termination"], exp["termination"])
    sol = assemble_solution(p, exp["positions"])
    data = wire.encode_solution(sol.image_names, sol.fact, sol.img_ptr, sol.feature_idx, sol.di, sol.dj)
    with open(os.path.join(GOLD, "%s_solution.pb" % name), "rb") as fh:
        assert data == fh.read()


def test_linesearch_fixture_exercises_contractions():
    _, exp = load("linesearch")
    assert int(exp["line_search_steps"]) > 20


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_reproduces_golden(b200, name):
    p, exp = load(name)
    pos, st = b200.solve(p)
    assert np.abs(pos - exp["positions"]).max() <= TOL_UNITS
    assert np.array_equal(st["iterations"], exp["iterations"])
    assert np.array_equal(st["termination"], exp["termination"])
    np.testing.assert_allclose(st["final_cost"], exp["final_cost"], rtol=1e-8, atol=1e-14)
