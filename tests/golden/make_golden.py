"""Generate the golden fixtures under tests/golden/.

The reference ships no golden vectors.  `*_solution.pb` is therefore produced by running the
REFERENCE'S OWN main() here: oracle/_ref/solve = multi-view-refinement/solve.cc + cost.cc + graph.cc
compiled unmodified against the shim headers of oracle/ref_shims/ (oracle/build_ref.py; Ceres'
minimizer restated in mini_ceres.cc, the Graclus cut replaced by csrc/lfr_cut.h — see DESIGN.md for
what that does and does not pin).  `*_expected.npz` holds the restated oracle's per-component
outputs (oracle/lfr_oracle.cc, literal line-search mode) for the same input; the script asserts
that the oracle pipeline reproduces the reference binary's SolutionFile byte for byte.

    python tests/golden/make_golden.py

writes, for each case, the MatchingFile bytes (`*_matches.pb`), the reference binary's SolutionFile
(`*_solution.pb`) and stdout (`*_stdout.txt`, timing lines removed), and the oracle's outputs
(`*_expected.npz`: positions, per-component iterations, termination codes, costs).
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


def run_reference_binary(matches_path, out_path):
    """oracle/_ref/solve on one input; returns its stdout without the timing lines."""
    import importlib.util
    import subprocess
    spec = importlib.util.spec_from_file_location("lfr_build_ref", os.path.join(ROOT, "oracle", "build_ref.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    exe = mod.build_solve()
    r = subprocess.run([exe, "--matches_file", matches_path, "--output_file", out_path], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return "".join(l + "\n" for l in r.stdout.splitlines() if " time:" not in l)


def main():
    orc = load_oracle()
    for name, ms in cases():
        data = wire.encode_matching_file(ms)
        mpath = os.path.join(HERE, "%s_matches.pb" % name)
        with open(mpath, "wb") as fh:
            fh.write(data)
        spath = os.path.join(HERE, "%s_solution.pb" % name)
        stdout = run_reference_binary(mpath, spath)          # the reference's own main()
        with open(os.path.join(HERE, "%s_stdout.txt" % name), "w") as fh:
            fh.write(stdout)
        p = build_problem(wire.decode_matching_file(data))
        pos, st = orc.solve(p, orc.default_options(n_threads=1))
        sol = assemble_solution(p, pos)
        mine = wire.encode_solution(sol.image_names, sol.fact, sol.img_ptr, sol.feature_idx, sol.di, sol.dj)
        with open(spath, "rb") as fh:
            assert fh.read() == mine, "%s: the oracle pipeline and the reference binary disagree" % name
        np.savez_compressed(os.path.join(HERE, "%s_expected.npz" % name), positions=pos,
                            iterations=st["iterations"], termination=st["termination"],
                            initial_cost=st["initial_cost"], final_cost=st["final_cost"],
                            track=p.track, comp=p.comp, is_root=p.is_root, comp_ptr=p.comp_ptr,
                            comp_nodes=p.comp_nodes, line_search_steps=st["total_line_search_steps"])
        print(name, "matches", ms.n_matches, "bytes", len(data), "iters", st["total_iterations"],
              "ls", st["total_line_search_steps"])


if __name__ == "__main__":
    main()
