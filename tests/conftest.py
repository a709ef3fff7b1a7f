import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle_util import load_oracle
    return load_oracle()


@pytest.fixture(scope="session")
def b200():
    """The product library; GPU tests fail loudly if it is missing."""
    from lfr_b200.capi import load_b200
    return load_b200()


_problem_cache = {}


def get_problem(name, scale=1.0, seed=None):
    key = (name, scale, seed)
    if key not in _problem_cache:
        from lfr_b200 import synth, build_problem
        ms = synth.generate(name, scale=scale, seed=seed)
        _problem_cache[key] = (ms, build_problem(ms))
    return _problem_cache[key]
