import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # the C-ABI library is a build artefact (git-ignored): build it when a fresh checkout has none
    # and nvcc is around (cross-compiles without a GPU); on the GPU box the shipped .so is used as is
    so = os.path.join(ROOT, "local-feature-refinement_b200", "csrc", "liblfr_b200.so")
    if not os.path.exists(so):
        import importlib.util
        import shutil
        if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"):
            spec = importlib.util.spec_from_file_location(
                "lfr_csrc_build", os.path.join(ROOT, "local-feature-refinement_b200", "csrc", "build.py"))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            mod.build()


def pytest_collection_modifyitems(config, items):
    """`gpu`-marked tests need a CUDA device: skip them (instead of failing on cudaGetDeviceCount) when a
    plain `pytest` runs on a CPU-only box.  `-m gpu` on the GPU box is unaffected."""
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    if have_gpu:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (the solve has no CPU fallback)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


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
