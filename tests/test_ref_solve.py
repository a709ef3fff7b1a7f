"""The whole pipeline against the REFERENCE'S OWN main().

oracle/_ref/solve is multi-view-refinement/solve.cc (+ cost.cc, graph.cc) compiled unmodified against
the shim headers of oracle/ref_shims/ (oracle/build_ref.py).  What runs is the reference author's
code for: protobuf reading incl. `.part.N` (solve.cc:412-481), node interning and edge lists (:53-65,
:474-478), constrained Kruskal (:487-549), root selection (:551-582), separate_meta_graph /
recursive_graph_cut / bfs (:162-373), create_and_solve_problem (:79-160: which residual blocks,
which loss, constant roots, bounds, solver options), the thread-pool dispatch (:599-635), the
output assembly (:643-679), the command line (:379-403) and every stdout line.  Behind the shims:
Ceres' minimizer is the mini-Ceres restatement, the Graclus cut is csrc/lfr_cut.h (the two pieces no
build in this image can pin), protobuf / Boost.PO / ThreadPool are functional equivalents.

Rows pinned here: A4, A5, A7, A8, H1-H5 and the drop-in boundary (b).
"""
import importlib.util
import os
import subprocess
import sys

import numpy as np
import pytest

from lfr_b200 import build_problem, synth, wire
from lfr_b200.solver import assemble_solution

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
LAUNCHER = os.path.join(ROOT, "multi-view-refinement", "build", "solve")


def ref_exe():
    spec = importlib.util.spec_from_file_location("lfr_build_ref", os.path.join(ROOT, "oracle", "build_ref.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    try:
        return mod.build_solve()
    except Exception:
        pytest.skip("oracle/_ref/solve is not built and the reference sources are not available here")


def run_ref(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([ref_exe()] + args, capture_output=True, text=True, env=e)


def oracle_pipeline(oracle, data, banned=()):
    """host stage (native) + oracle solve + output assembly -> (SolutionFile bytes, problem, stats, log lines)"""
    lines = []
    p = build_problem(wire.decode_matching_file(data), banned_images=banned, log=lines.append)
    pos, st = oracle.solve(p, oracle.default_options(n_threads=4))
    sol = assemble_solution(p, pos)
    out = wire.encode_solution(sol.image_names, sol.fact, sol.img_ptr, sol.feature_idx, sol.di, sol.dj)
    lines.append("# points with at least one coordinate > 0.5: %d" % sol.n_outside)
    return out, p, st, lines


def untimed(stdout):
    return [l for l in stdout.splitlines() if " time:" not in l]


SCENES = [("cfg1", 1.0, None), ("cfg2", 0.1, 3), ("cfg2", 0.25, 11), ("cfg3", 0.05, 5), ("cfg4", 0.05, 7),
          ("ring60", 0.5, 2)]


@pytest.mark.parametrize("name,scale,seed", SCENES)
def test_solution_file_is_byte_identical_to_the_reference_binary(oracle, tmp_path, name, scale, seed):
    ms = synth.generate(name, scale=scale, seed=seed)
    data = wire.encode_matching_file(ms)
    mpath, opath, rpath = tmp_path / "m.pb", tmp_path / "s.pb", tmp_path / "rec.txt"
    mpath.write_bytes(data)
    r = run_ref(["--matches_file", str(mpath), "--output_file", str(opath), "--n_threads", "4"],
                env={"LFR_CERES_RECORD_FILE": str(rpath)})
    assert r.returncode == 0, r.stderr
    mine, p, st, lines = oracle_pipeline(oracle, data)
    assert opath.read_bytes() == mine
    # every stdout line except the three timings (solve.cc:484-485,534,549,591,606,670)
    assert untimed(r.stdout) == [l for l in lines if " time:" not in l]
    # per-problem LM iteration counts: ceres::Solve summaries recorded by the shim, matched to the
    # dispatch slots through the lowest parameter-block address (= lowest node index of the component)
    rec = np.loadtxt(str(rpath)).reshape(-1, 9)
    sizes = np.diff(p.comp_ptr.astype(np.int64))
    solved = np.nonzero(sizes > 1)[0]
    assert rec.shape[0] == solved.shape[0]
    min_node = np.array([p.comp_nodes[p.comp_ptr[c]] for c in solved])       # nodes ascending inside a component
    order_mine = solved[np.argsort(min_node)]
    order_ref = np.argsort(rec[:, 0])
    assert np.array_equal(rec[order_ref, 1].astype(np.int64), st["iterations"][order_mine])
    assert int(rec[:, 3].sum()) == int(st["total_line_search_steps"])
    # Summary::initial_cost includes the fixed cost of the all-constant (root-root) blocks; the C ABI reports it without
    np.testing.assert_allclose(rec[order_ref, 4] - rec[order_ref, 8], st["initial_cost"][order_mine], rtol=1e-11, atol=1e-14)


@pytest.mark.parametrize("name", ["tiny", "linesearch", "fountain_2pct"])
def test_committed_goldens_are_the_reference_binarys_output(tmp_path, name):
    opath = tmp_path / "s.pb"
    r = run_ref(["--matches_file", os.path.join(GOLD, name + "_matches.pb"), "--output_file", str(opath)])
    assert r.returncode == 0
    assert opath.read_bytes() == open(os.path.join(GOLD, name + "_solution.pb"), "rb").read()
    assert untimed(r.stdout) == open(os.path.join(GOLD, name + "_stdout.txt")).read().splitlines()


def test_banned_images_and_part_files(oracle, tmp_path):
    ms = synth.generate("cfg2", scale=0.08, seed=21)
    data = wire.encode_matching_file(ms)
    banned = [ms.image_names[2], ms.image_names[5]]
    m = tmp_path / "m.pb"
    m.write_bytes(data)
    o = tmp_path / "o.pb"
    r = run_ref(["--matches_file", str(m), "--output_file", str(o), "--banned_images", banned[0],
                 "--banned_images=" + banned[1]])
    assert r.returncode == 0
    mine, _, _, lines = oracle_pipeline(oracle, data, banned=banned)
    assert o.read_bytes() == mine and untimed(r.stdout) == [l for l in lines if " time:" not in l]
    # matches split over .part.N files (compute_match_graph.py:189-205, solve.cc:416-424)
    half = ms.n_pairs // 2
    parts = [ms.select_pairs(np.arange(0, half)), ms.select_pairs(np.arange(half, ms.n_pairs))] \
        if hasattr(ms, "select_pairs") else None
    if parts is None:
        pytest.skip("MatchSet.select_pairs not available")
    base = tmp_path / "split.pb"
    for k, part in enumerate(parts):
        (tmp_path / ("split.pb.part.%d" % k)).write_bytes(wire.encode_matching_file(part))
    o2 = tmp_path / "o2.pb"
    r = run_ref(["--matches_file", str(base), "--output_file", str(o2)])
    assert r.returncode == 0
    mine2, _, _, _ = oracle_pipeline(oracle, data)
    assert o2.read_bytes() == mine2


BAD_COMMAND_LINES = [["--help"], ["--matches_file", "x"], ["--foo", "1"], ["pos"], ["--matches_file"],
                     ["--matches_file", "a", "--matches_file", "b", "--output_file", "c"],
                     ["--n_threads", "abc", "--matches_file", "a", "--output_file", "b"],
                     ["--mat", "x", "--out", "y", "-x"], ["--output_file=y"]]


@pytest.mark.parametrize("argv", BAD_COMMAND_LINES)
def test_command_line_errors_match_the_reference_binary(argv):
    """solve.cc:379-403 (Boost.Program_options): same exit code, stdout and stderr from the drop-in
    launcher (argument handling needs no GPU)."""
    a = run_ref(argv)
    b = subprocess.run([sys.executable, LAUNCHER] + argv, capture_output=True, text=True)
    assert (a.returncode, a.stdout, a.stderr) == (b.returncode, b.stdout, b.stderr)


def test_parse_failure_exit_code(tmp_path):
    bad = tmp_path / "bad.pb"
    bad.write_bytes(b"\x0a\xff\xff\xff\xff\x0f not a protobuf")
    r = run_ref(["--matches_file", str(bad), "--output_file", str(tmp_path / "o.pb")])
    assert r.returncode == 255 and "Failed to parse proto object." in r.stderr      # return -1, solve.cc:433-436


@pytest.mark.gpu
@pytest.mark.parametrize("host", ["native", "python"])
@pytest.mark.parametrize("name", ["tiny", "linesearch", "fountain_2pct"])
def test_gpu_launcher_reproduces_the_reference_binarys_solution_file(tmp_path, name, host):
    """The drop-in executable on the GPU against the reference's own main(): same stdout lines, and
    the SolutionFile equal float for float (fp32 on the wire) — byte-identical unless a displacement
    sits within 1e-16 of a float32 rounding boundary."""
    m = os.path.join(GOLD, name + "_matches.pb")
    o = tmp_path / "gpu.pb"
    env = dict(os.environ)
    if host == "python":
        env["LFR_SOLVE_PYTHON"] = "1"     # the launcher hands over to solve_native otherwise
    r = subprocess.run([sys.executable, LAUNCHER, "--matches_file", m, "--output_file", str(o)],
                       capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    want = open(os.path.join(GOLD, name + "_solution.pb"), "rb").read()
    assert untimed(r.stdout) == open(os.path.join(GOLD, name + "_stdout.txt")).read().splitlines()
    got = o.read_bytes()
    if got != want:
        a, b = wire.decode_solution(got), wire.decode_solution(want)    # lists of (name, fact, feature_idx, di, dj)
        assert [x[0] for x in a] == [x[0] for x in b]
        n_diff = 0
        for (_, fa, ia, dia, dja), (_, fb, ib, dib, djb) in zip(a, b):
            assert fa == fb and np.array_equal(ia, ib)
            assert np.abs(dia - dib).max() <= 1e-4 / 16 and np.abs(dja - djb).max() <= 1e-4 / 16
            n_diff += int((dia != dib).sum() + (dja != djb).sum())
        assert n_diff <= 2
