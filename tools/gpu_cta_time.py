import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import torch
from lfr_b200 import build_problem, synth
from lfr_b200.capi import Plan, load_b200
lib = load_b200()
for name in sys.argv[1:] or ["ring200"]:
    p = build_problem(synth.generate(name))
    plan = Plan(lib, p)
    s = torch.cuda.current_stream().cuda_stream
    plan.solve(s); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); plan.solve(s); b.record(); torch.cuda.synchronize()
    pos, st = plan.download(s)
    print(name, "solve ms %.3f" % a.elapsed_time(b), "lm iters", st["total_iterations"], "launches", plan.num_launches())
