import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import ctypes as C
import torch
import bench
from lfr_b200.capi import load_b200, Plan
lib = load_b200()
p, n_tracks = bench.build_workload(sys.argv[1] if len(sys.argv) > 1 else "cfg4")
s2, keep, pos_pinned, h2d = bench.pinned_problem(lib, p)
opts = lib.default_options()
stt, bufs = lib.make_stats(p.n_components)
for i in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    rc = lib.lib.lfr_solve(C.byref(s2), C.byref(opts), pos_pinned.data_ptr(), C.byref(stt))
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("lfr_solve wall %.3f ms  h2d %.3f kern %.3f d2h %.3f  launches %d" % (dt * 1e3, stt.h2d_ms, stt.kernel_ms, stt.d2h_ms, stt.n_kernel_launches))
plan = Plan(lib, p)
for i in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); plan.solve(); torch.cuda.synchronize(); t1 = time.perf_counter()
    pos, st = plan.download(); t2 = time.perf_counter()
    print("plan solve %.3f ms  download %.3f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3))
