"""One solve per tier, meant to run under compute-sanitizer:
    compute-sanitizer --tool memcheck  python tools/gpu_sanitize.py
    compute-sanitizer --tool racecheck python tools/gpu_sanitize.py
    compute-sanitizer --tool initcheck python tools/gpu_sanitize.py
Small scenes so that the 10-100x slowdown stays within a minute."""
import os
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np  # noqa: E402

from lfr_b200 import build_problem, synth  # noqa: E402
from lfr_b200.capi import load_b200  # noqa: E402

lib = load_b200()
which = sys.argv[1:] or ["cfg1", "warp", "tile", "ring", "pcg", "v1"]
for w in which:
    env = None
    if w == "cfg1":
        ms = synth.generate("cfg1")
        opts = lib.default_options()
    elif w == "warp":      # register tier up to n ~ 22
        ms = synth.generate("cfg2", scale=0.15)
        opts = lib.default_options()
    elif w == "tile":      # components with 32 < n <= 80 unknowns
        ms = synth.generate("cfg4", scale=0.1)
        opts = lib.default_options()
    elif w == "ring":      # CTA tier (n > 96)
        ms = synth.generate("ring60", scale=0.25)
        opts = lib.default_options()
    elif w == "pcg":       # every component through the CTA tier
        ms = synth.generate("cfg1")
        opts = lib.default_options(linear_solver=2)
    elif w == "v1":
        ms = synth.generate("cfg1")
        opts = lib.default_options(debug_flags=1)  # LFR_DBG_FORCE_SMEM_CHOLESKY
    p = build_problem(ms)
    pos, st = lib.solve(p, opts)
    sizes = np.diff(p.comp_ptr.astype(np.int64))
    print(w, "components", p.n_components, "max nodes", int(sizes.max()) if sizes.size else 0,
          "iterations", int(st["total_iterations"]), "finite", bool(np.isfinite(pos).all()))
