"""CTA / PCG tier against the exact-Cholesky oracle: per-component agreement on ring200, ring60 seeds
and a Madrid-topology scene (cfg5 at 5 % keypoints: components of up to 1000 nodes = 2000 unknowns)."""
import json
import os
import sys
import time

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np  # noqa: E402

from lfr_b200 import build_problem, synth  # noqa: E402
from lfr_b200.capi import load_b200  # noqa: E402
from oracle_util import load_oracle  # noqa: E402

lib, orc = load_b200(), load_oracle()
TOL = 1e-4 / 16
out = {}
for name, scale, seed in [("ring200", 1.0, None), ("ring60", 1.0, None), ("ring60", 0.3, 101), ("ring60", 0.3, 202),
                          ("ring60", 0.3, 303), ("ring200", 0.5, 7), ("cfg5", 0.05, None)]:
    p = build_problem(synth.generate(name, scale=scale, seed=seed))
    t0 = time.time()
    pos_g, st_g = lib.solve(p)
    t1 = time.time()
    pos_o, st_o = orc.solve(p, orc.default_options(n_threads=os.cpu_count()))
    t2 = time.time()
    sizes = np.diff(p.comp_ptr.astype(np.int64))
    err = np.zeros(p.n_components)
    for c in range(p.n_components):
        nodes = p.comp_nodes[p.comp_ptr[c]:p.comp_ptr[c + 1]].astype(int)
        err[c] = np.abs(pos_g[nodes] - pos_o[nodes]).max()
    same_it = st_g["iterations"] == st_o["iterations"]
    good = (err <= TOL) & same_it
    solved = sizes > 1
    bad = np.nonzero(solved & ~good)[0]
    rec = {"components": int(solved.sum()), "max_nodes": int(sizes.max()), "good": int(good[solved].sum()),
           "max_err_units": float(err.max()), "max_err_px": float(err.max() * 16), "gpu_s": t1 - t0, "oracle_s": t2 - t1,
           "kernel_ms": st_g["kernel_ms"], "lm_iterations": int(st_g["total_iterations"]),
           "bad": [(int(c), int(sizes[c]), float(err[c]), int(st_g["iterations"][c]), int(st_o["iterations"][c]),
                    float(st_g["final_cost"][c]), float(st_o["final_cost"][c])) for c in bad[:10]]}
    out["%s_%s_%s" % (name, scale, seed)] = rec
    print(name, scale, seed, json.dumps(rec), flush=True)
os.makedirs(os.path.join(R, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(R, "gpurun_out", "r2e_cta_parity.json"), "w"), indent=1)
