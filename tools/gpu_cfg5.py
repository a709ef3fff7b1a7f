"""Madrid-Metropolis-scale run (BASELINE.json configs[4]) on one GPU: timing and
size-independent properties (the dense CPU oracle cannot check 2000-unknown
components in reasonable time; parity of the CTA tier is tested on ring60 /
ring200)."""
import json
import os
import sys
import time

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np  # noqa: E402

from lfr_b200 import build_problem, refined_track_count, synth  # noqa: E402
from lfr_b200.capi import Plan, load_b200  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
t0 = time.time()
ms = synth.generate(name)
t1 = time.time()
p = build_problem(ms)
t2 = time.time()
lib = load_b200()
plan = Plan(lib, p)
t3 = time.time()
import torch  # noqa: E402

s = torch.cuda.current_stream().cuda_stream
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
plan.solve(s)
torch.cuda.synchronize()
e0.record()
plan.solve(s)
e1.record()
torch.cuda.synchronize()
ms_solve = e0.elapsed_time(e1)
pos, st = plan.download(s)
sizes = np.diff(p.comp_ptr.astype(np.int64))
solved = sizes > 1
out = {
    "workload": name, "matches": int(ms.n_matches), "nodes": int(p.graph.n_nodes), "directed_edges": int(p.graph.n_edges),
    "tracks": int(p.info["n_tracks"]), "components": int(p.n_components), "max_component_nodes": int(sizes.max()),
    "components_over_48_nodes": int((sizes > 48).sum()),
    "gen_s": t1 - t0, "host_stage_s": t2 - t1, "plan_create_s": t3 - t2, "solve_ms": ms_solve,
    "launches": plan.num_launches(), "lm_iterations": int(st["total_iterations"]),
    "line_search_steps": int(st["total_line_search_steps"]),
    "tracks_refined": refined_track_count(p),
    "tracks_per_s": refined_track_count(p) / (ms_solve / 1e3),
    "termination_hist": np.bincount(st["termination"], minlength=8).tolist(),
    "max_iterations": int(st["iterations"].max()),
    # size-independent properties
    "cost_never_increases": bool(np.all(st["final_cost"][solved] <= st["initial_cost"][solved] * (1 + 1e-12) + 1e-15)),
    "within_bounds": bool(np.abs(pos).max() <= 1.0),
    "roots_untouched": bool(np.all(pos[p.is_root.astype(bool)] == 0)),
    "all_finite": bool(np.isfinite(pos).all()),
    "cost_initial": float(st["initial_cost"].sum()), "cost_final": float(st["final_cost"].sum()),
}
plan.solve(s)
pos2, _ = plan.download(s)
out["deterministic"] = bool(np.array_equal(pos, pos2))
print(json.dumps(out, indent=1))
os.makedirs(os.path.join(R, "gpurun_out"), exist_ok=True)
with open(os.path.join(R, "gpurun_out", "%s_run.json" % name), "w") as fh:
    json.dump(out, fh, indent=1)
