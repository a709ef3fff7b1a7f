import sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lfr_b200.capi import load_b200, Plan
from lfr_b200 import synth, build_problem
lib = load_b200()
cfg = sys.argv[1] if len(sys.argv) > 1 else 'cfg2'
ms = synth.generate(cfg); p = build_problem(ms)
plan = Plan(lib, p)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1
for _ in range(n):
    plan.solve()
pos, st = plan.download()
print('iters', st['total_iterations'])
