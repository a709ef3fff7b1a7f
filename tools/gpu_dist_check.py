"""Sharded multi-GPU solve through the CLI (`solve --gpus N`, NCCL) against the 1-GPU result."""
import os, subprocess, sys, tempfile
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
from lfr_b200 import synth, wire
name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
d = tempfile.mkdtemp()
m = os.path.join(d, "m.pb")
wire.write_matching_file(synth.generate(name), m, pairs_per_part=5000)
solve = os.path.join(R, "multi-view-refinement", "build", "solve")
outs = []
for g in (1, n):
    o = os.path.join(d, "s%d.pb" % g)
    r = subprocess.run([solve, "--matches_file", m, "--output_file", o, "--gpus", str(g), "--stats_json", o + ".json"],
                       capture_output=True, text=True, cwd=R)
    print("gpus", g, "rc", r.returncode, [l for l in r.stdout.splitlines() if "time" in l.lower()], r.stderr[-300:] if r.returncode else "")
    outs.append(open(o, "rb").read())
print("bitwise identical solution files:", outs[0] == outs[1], len(outs[0]))
