"""Command-line surface of the `solve` drop-in (solve.cc:375-682)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from lfr_b200 import cli, synth, wire

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SOLVE = os.path.join(ROOT, "multi-view-refinement", "build", "solve")


def test_launcher_exists_where_benchmark_py_expects_it():
    assert os.path.exists(SOLVE) and os.access(SOLVE, os.X_OK)


def test_help_and_argument_errors():
    r = subprocess.run([sys.executable, SOLVE, "--help"], capture_output=True, text=True)
    assert r.returncode == 0 and "Patch Match graph problem solver" in r.stdout and "--matches_file" in r.stdout
    r = subprocess.run([sys.executable, SOLVE, "--output_file", "x"], capture_output=True, text=True)
    assert r.returncode == 1 and "ERROR:" in r.stderr and "matches_file" in r.stderr      # solve.cc:397-401
    r = subprocess.run([sys.executable, SOLVE, "--matches_file", "a", "--output_file", "b", "--bogus", "1"],
                       capture_output=True, text=True)
    assert r.returncode == 1 and "ERROR:" in r.stderr


def test_unparsable_input_returns_minus_one(tmp_path):
    bad = tmp_path / "bad.pb"
    bad.write_bytes(b"\x0a\xff\xff\xff\xff\x0f")
    r = subprocess.run([sys.executable, SOLVE, "--matches_file", str(bad), "--output_file", str(tmp_path / "o.pb")],
                       capture_output=True, text=True)
    assert r.returncode == 255 and "Failed to parse proto object." in r.stderr            # solve.cc:433-436


def _run_main_with_oracle(monkeypatch, oracle, argv, capsys):
    """Drive cli.main in-process with the GPU solve swapped for the oracle
    (CPU test of everything around the solve)."""
    import lfr_b200.solver as solver
    monkeypatch.setattr(solver, "solve_problem", lambda p, options=None, positions=None: oracle.solve(p))
    rc = cli.main(argv)
    return rc, capsys.readouterr().out


def test_cli_end_to_end_with_injected_solver(tmp_path, monkeypatch, oracle, capsys):
    ms = synth.generate("cfg1")
    mpath, opath, spath = str(tmp_path / "m.pb"), str(tmp_path / "s.pb"), str(tmp_path / "st.json")
    wire.write_matching_file(ms, mpath, pairs_per_part=2)          # parts only, like large scenes
    rc, out = _run_main_with_oracle(monkeypatch, oracle,
                                    ["--matches_file", mpath, "--output_file", opath, "--stats_json", spath], capsys)
    assert rc == 0
    for line in ("# graph nodes:", "# graph edges:", "# tracks:", "max track size:", "Graph-cut time:",
                 "# components:", "max component size:", "Solver time:", "Total time:",
                 "# points with at least one coordinate > 0.5:"):
        assert line in out, line
    from lfr_b200 import build_problem
    p = build_problem(ms)
    pos, _ = oracle.solve(p)
    sol = wire.decode_solution(open(opath, "rb").read())
    # images in first-node-appearance order, every node present, values = float32(positions)
    assert [s[0] for s in sol] == ["0000.png", "0001.png", "0002.png"]
    assert sum(len(s[2]) for s in sol) == p.graph.n_nodes
    for name, fact, fi, di, dj in sol:
        img = ms.image_names.index(name)
        nodes = np.nonzero(p.graph.node_image == img)[0]
        assert fact == 1.0 and np.array_equal(fi, p.graph.node_feat[nodes])
        assert np.array_equal(di, pos[nodes, 0].astype(np.float32)) and np.array_equal(dj, pos[nodes, 1].astype(np.float32))
    import json
    st = json.load(open(spath))
    assert st["tracks_refined"] > 0 and st["lm_iterations"] > 0
    # banned images drop pairs before the graph is built (solve.cc:444-446)
    rc, out = _run_main_with_oracle(monkeypatch, oracle, ["--matches_file", mpath, "--output_file", opath,
                                                          "--banned_images", "0002.png"], capsys)
    assert rc == 0 and [s[0] for s in wire.decode_solution(open(opath, "rb").read())] == ["0000.png", "0001.png"]


def test_missing_matches_file_writes_empty_solution(tmp_path, monkeypatch, oracle, capsys):
    """No file and no parts: the reference parses nothing and writes an empty SolutionFile."""
    opath = str(tmp_path / "s.pb")
    rc, out = _run_main_with_oracle(monkeypatch, oracle, ["--matches_file", str(tmp_path / "nope.pb"),
                                                          "--output_file", opath], capsys)
    assert rc == 0 and os.path.getsize(opath) == 0 and "# graph nodes: 0" in out


@pytest.mark.gpu
def test_launcher_on_gpu_matches_oracle(tmp_path, oracle):
    ms = synth.generate("cfg1")
    mpath, opath = str(tmp_path / "m.pb"), str(tmp_path / "s.pb")
    wire.write_matching_file(ms, mpath)
    r = subprocess.run([SOLVE, "--matches_file", mpath, "--output_file", opath], capture_output=True, text=True,
                       cwd=ROOT)
    assert r.returncode == 0, r.stderr
    assert "Solver time:" in r.stdout
    from lfr_b200 import build_problem
    p = build_problem(ms)
    pos, _ = oracle.solve(p)
    sol = wire.decode_solution(open(opath, "rb").read())
    got = np.concatenate([np.stack([s[3], s[4]], 1) for s in sol])
    order = np.argsort(np.argsort(p.graph.node_image, kind="stable"), kind="stable")
    want = pos[np.argsort(p.graph.node_image, kind="stable")]
    assert np.abs(got - want.astype(np.float32)).max() <= 1e-4 / 16
