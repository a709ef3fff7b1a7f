import sys, time
import os; R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R,'tests'))
import numpy as np
from oracle_util import load_oracle
from lfr_b200.capi import load_b200, Plan
from lfr_b200 import synth, build_problem
orc = load_oracle(); lib = load_b200()
for cfg in ['cfg1', 'cfg2']:
    ms = synth.generate(cfg); p = build_problem(ms)
    pos_o, st_o = orc.solve(p, orc.default_options(n_threads=8))
    t = time.time(); pos_g, st_g = lib.solve(p); dt = time.time() - t
    print(cfg, 'err', np.abs(pos_g - pos_o).max(), 'iters', st_g['total_iterations'], st_o['total_iterations'],
          'term eq', np.array_equal(st_g['termination'], st_o['termination']), 'iter eq', np.array_equal(st_g['iterations'], st_o['iterations']),
          'h2d/k/d2h ms', st_g['h2d_ms'], st_g['kernel_ms'], st_g['d2h_ms'], 'wall', dt*1e3, 'cpu ms', st_o['total_ms'])
    bad = np.nonzero(st_g['iterations'] != st_o['iterations'])[0]
    print(' mismatching comps', bad[:10], st_g['iterations'][bad[:10]], st_o['iterations'][bad[:10]])
    plan = Plan(lib, p)
    import torch
    for _ in range(3): plan.solve()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    s = torch.cuda.current_stream().cuda_stream
    ev[0].record(); 
    for _ in range(10): plan.solve(s)
    ev[1].record(); torch.cuda.synchronize()
    print(' plan solve ms', ev[0].elapsed_time(ev[1]) / 10, 'launches', plan.num_launches(), 'traffic', plan.traffic(s))
