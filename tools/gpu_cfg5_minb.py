"""cfg5 (or another scene) on one GPU: CTA-tier occupancy settings (LFR_CTA_MINB = register caps of
the four size classes) against the solve time; results must be bitwise identical across settings."""
import hashlib
import os
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np  # noqa: E402

from lfr_b200 import build_problem, synth  # noqa: E402
from lfr_b200.capi import Plan, load_b200  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
settings = sys.argv[2:] or ["2,2,2,2", "2,2,3,3", "2,2,3,4", "2,2,4,4", "2,2,2,4"]
p = build_problem(synth.generate(name))
lib = load_b200()
import torch  # noqa: E402

s = torch.cuda.current_stream().cuda_stream
for st in settings:
    parts = (st.split("/") + ["", ""])[:3]
    os.environ["LFR_CTA_MINB"] = parts[0]
    os.environ["LFR_CTA_SMEM_VECS"] = parts[1] or "6,6,6,6"
    os.environ["LFR_CTA_NT"] = parts[2] or "256,256,256,256"
    plan = Plan(lib, p)
    ts = []
    for i in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        plan.solve(s)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    pos, stt = plan.download(s)
    print("minb %s  launches %d  solve ms %s  iterations %d  sha %s" % (
        st, plan.num_launches(), " ".join("%.1f" % t for t in ts), int(stt["total_iterations"]),
        hashlib.sha1(pos.tobytes()).hexdigest()[:12]), flush=True)
    del plan
