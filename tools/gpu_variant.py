"""A/B of a library variant built with csrc/build.py::build_variant: device-resident solve time on
cfg2 / cfg3 and 16-scene throughput.   python tools/gpu_variant.py <lib.so> [<lib2.so> ...]"""
import os
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from lfr_b200 import build_problem, synth  # noqa: E402
from lfr_b200 import capi  # noqa: E402
from lfr_b200.matchset import MatchSet  # noqa: E402

probs = {}
for name in ("cfg2", "cfg3"):
    probs[name] = build_problem(synth.generate(name))
scenes = []


flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for path in sys.argv[1:]:
    lib = capi.Library(os.path.join(R, "local-feature-refinement_b200", "csrc", path) if not os.path.isabs(path) else path)
    for name, p in probs.items():
        plan = capi.Plan(lib, p)
        s = torch.cuda.current_stream().cuda_stream
        for _ in range(3):
            plan.solve(s)
        torch.cuda.synchronize()
        ts = []
        for _ in range(15):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            plan.solve(s)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        print("%-28s %-10s median %.4f ms  min %.4f ms" % (os.path.basename(path), name, float(np.median(ts)), min(ts)), flush=True)
        plan.close()
