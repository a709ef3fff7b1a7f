"""`multi-view-refinement/build/solve` — command-line surface of solve.cc:375-682.

    solve --matches_file X.pb --output_file Y.pb [--n_threads 8]
          [--banned_images NAME]...                      (solve.cc:379-385)

Same flags, same `.part.N` handling (solve.cc:416-424), same stdout lines
(solve.cc:484-485,534,549,589,591,606,638,641,670) and exit codes (0; 1 on a
command-line error, solve.cc:397-401; -1 = 255 when the input does not parse or
the output cannot be written, solve.cc:433-436,674-677), so
local-feature-evaluation/benchmark.py:100-104, eth/benchmark.py:108-112 and
custom_demo.py:101-105 run unchanged.  `--n_threads` is accepted and ignored
(the solve runs on the GPU).  Opt-in extras: --device, --gpus, --stats_json.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

USAGE = """Options:
  --help                    print the help
  --matches_file arg        path to the matches file
  --output_file arg         path to the output file
  --n_threads arg (=8)      # threads
  --banned_images arg (={}) banned images
"""


class _Parser(argparse.ArgumentParser):
    def error(self, message):  # solve.cc:397-401
        sys.stderr.write("ERROR: %s\n\n" % message)
        sys.stderr.write(USAGE)
        raise SystemExit(1)


def parse_args(argv):
    ap = _Parser(prog="solve", add_help=False)
    ap.add_argument("--help", action="store_true")
    ap.add_argument("--matches_file")
    ap.add_argument("--output_file")
    ap.add_argument("--n_threads", type=int, default=8)
    ap.add_argument("--banned_images", action="append", default=[])
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--stats_json", default=None)
    args = ap.parse_args(argv)
    if args.help:
        sys.stdout.write("Patch Match graph problem solver\n\n" + USAGE)
        raise SystemExit(0)
    if args.matches_file is None:
        ap.error("the option '--matches_file' is required but missing")
    if args.output_file is None:
        ap.error("the option '--output_file' is required but missing")
    return args


def main(argv=None) -> int:
    args = parse_args(sys.argv[1:] if argv is None else argv)
    if args.gpus > 1 and "RANK" not in os.environ:
        # one process per GPU: re-launch under torch.distributed.run (NCCL over NVLink)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 2000),
               os.path.abspath(__file__)] + list(sys.argv[1:] if argv is None else argv)
        return subprocess.call(cmd)

    from . import wire
    from .solver import refine
    from .capi import load_b200
    from .graph import refined_track_count

    rank = int(os.environ.get("RANK", "0"))
    say = (lambda s: print(s, flush=True)) if rank == 0 else (lambda s: None)
    if not wire.matches_files(args.matches_file):
        # the reference parses zero files and writes an empty solution
        pass
    try:
        if args.matches_file.endswith(".npz") and os.path.exists(args.matches_file):
            from .matchset import MatchSet
            ms = MatchSet.load_npz(args.matches_file)   # packed arrays written by MatchSet.save_npz (opt-in extra)
        else:
            ms = wire.read_matching_file(args.matches_file) if wire.matches_files(args.matches_file) else None
    except wire.ParseError:
        sys.stderr.write("Failed to parse proto object.\n")
        return 255
    if ms is None:
        from .matchset import MatchSet
        import numpy as np
        ms = MatchSet([], np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0, np.float32),
                      np.zeros(0, np.float32), np.zeros(1, np.int64), np.zeros(0, np.uint32),
                      np.zeros(0, np.uint32), np.zeros(0, np.float32), np.zeros((0, 18), np.float32),
                      np.zeros((0, 18), np.float32))

    lib = load_b200()   # fails loudly when the CUDA library is missing
    opts = lib.default_options(device=args.device, n_threads=args.n_threads)
    t_start = time.perf_counter()
    timing = {}

    def solve_fn(p):
        t1 = time.perf_counter()
        if "RANK" in os.environ and int(os.environ.get("WORLD_SIZE", "1")) > 1:
            from .dist import solve_distributed
            pos, st = solve_distributed(p)
        else:
            from .solver import solve_problem
            pos, st = solve_problem(p, opts)
        timing["solver_ms"] = (time.perf_counter() - t1) * 1e3
        return pos, st

    p, pos, st, sol = refine(ms, args.banned_images, opts, log=say, solve_fn=solve_fn)
    say("Solver time: %dms" % int(timing.get("solver_ms", 0.0)))                       # solve.cc:638
    say("Total time: %dms" % int((time.perf_counter() - t_start) * 1e3))               # solve.cc:641
    say("# points with at least one coordinate > 0.5: %d" % sol.n_outside)             # solve.cc:670
    if rank != 0:
        return 0
    try:
        data = wire.encode_solution(sol.image_names, sol.fact, sol.img_ptr, sol.feature_idx, sol.di, sol.dj)
        with open(args.output_file, "wb") as fh:
            fh.write(data)
    except OSError:
        sys.stderr.write("Failed to write proto object.\n")
        return 255
    if args.stats_json:
        n_tracks = refined_track_count(p)
        solver_s = max(timing.get("solver_ms", 0.0), 1e-9) / 1e3
        out = dict(n_nodes=p.graph.n_nodes, n_edges=p.graph.n_edges, info=p.info,
                   tracks_refined=n_tracks, lm_iterations=int(st.get("total_iterations", 0)),
                   solver_ms=timing.get("solver_ms"), kernel_ms=st.get("kernel_ms"), h2d_ms=st.get("h2d_ms"),
                   d2h_ms=st.get("d2h_ms"), tracks_per_s=n_tracks / solver_s,
                   lm_iters_per_s=int(st.get("total_iterations", 0)) / solver_s)
        with open(args.stats_json, "w") as fh:
            json.dump(out, fh, indent=1, default=float)
    return 0


if __name__ == "__main__":
    if __package__ in (None, ""):
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        import lfr_b200.cli as _cli
        raise SystemExit(_cli.main())
    raise SystemExit(main())
