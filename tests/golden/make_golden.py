"""Generate the golden fixtures under tests/golden/.

The reference ships no golden vectors and cannot be built or imported here
(C++ against Ceres/COLMAP), so these fixtures are produced by the
CPU oracle (oracle/lfr_oracle.cc) — they pin the oracle and the CUDA path
against silent drift, they do not pin either against a real Ceres binary.

    python tests/golden/make_golden.py

writes, for each case, the MatchingFile bytes (`*_matches.pb`) and the
oracle's outputs (`*_expected.npz`: positions, per-component iterations,
termination codes, costs; SolutionFile bytes `*_solution.pb`).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle_util import load_oracle  # noqa: E402
from lfr_b200 import build_problem, synth, wire  # noqa: E402
from lfr_b200.solver import assemble_solution  # noqa: E402


def cases():
    yield "tiny", synth.generate("cfg1", scale=0.3, seed=4242)
    ms = synth.generate("cfg1", scale=0.25, seed=99)
    rng = np.random.default_rng(5)
    ms.disp1[:] = rng.uniform(-1.2, 1.2, size=ms.disp1.shape).astype(np.float32)   # forces Armijo contractions
    ms.disp2[:] = rng.uniform(-1.2, 1.2, size=ms.disp2.shape).astype(np.float32)
    yield "linesearch", ms
    yield "fountain_2pct", synth.generate("cfg2", scale=0.02, seed=7)


def main():
    orc = load_oracle()
    for name, ms in cases():
        data = wire.encode_matching_file(ms)
        with open(os.path.join(HERE, "%s_matches.pb" % name), "wb") as fh:
            fh.write(data)
        p = build_problem(wire.decode_matching_file(data))
        pos, st = orc.solve(p, orc.default_options(n_threads=1))
        sol = assemble_solution(p, pos)
        with open(os.path.join(HERE, "%s_solution.pb" % name), "wb") as fh:
            fh.write(wire.encode_solution(sol.image_names, sol.fact, sol.img_ptr, sol.feature_idx, sol.di, sol.dj))
        np.savez_compressed(os.path.join(HERE, "%s_expected.npz" % name), positions=pos,
                            iterations=st["iterations"], termination=st["termination"],
                            initial_cost=st["initial_cost"], final_cost=st["final_cost"],
                            track=p.track, comp=p.comp, is_root=p.is_root, comp_ptr=p.comp_ptr,
                            comp_nodes=p.comp_nodes, line_search_steps=st["total_line_search_steps"])
        print(name, "matches", ms.n_matches, "bytes", len(data), "iters", st["total_iterations"],
              "ls", st["total_line_search_steps"])


if __name__ == "__main__":
    main()
