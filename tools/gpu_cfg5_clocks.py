"""cfg5 on one GPU: per-solve time of consecutive plan solves with SM clock / power samples alongside
(does the 0.4-0.5 s fp64-heavy solve run into the power cap?)."""
import os
import subprocess
import sys
import threading
import time

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np  # noqa: E402

from lfr_b200 import build_problem, synth  # noqa: E402
from lfr_b200.capi import Plan, load_b200  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
n_solves = int(sys.argv[2]) if len(sys.argv) > 2 else 12
p = build_problem(synth.generate(name))
lib = load_b200()
import torch  # noqa: E402

samples = []
stop = False


def sampler():
    while not stop:
        r = subprocess.run(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit,temperature.gpu,clocks_throttle_reasons.active",
                            "--format=csv,noheader,nounits"], capture_output=True, text=True)
        samples.append((time.perf_counter(), r.stdout.strip()))
        time.sleep(0.05)


plan = Plan(lib, p)
s = torch.cuda.current_stream().cuda_stream
th = threading.Thread(target=sampler, daemon=True)
th.start()
time.sleep(0.5)
marks = []
for i in range(n_solves):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    plan.solve(s)
    e1.record()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    marks.append((t0, t1, e0.elapsed_time(e1)))
    if i == n_solves // 2:
        time.sleep(3.0)   # let the device cool: does the next solve get faster again?
stop = True
th.join()
for i, (t0, t1, ms) in enumerate(marks):
    inside = [x for (t, x) in samples if t0 <= t <= t1]
    print("solve %2d  %.1f ms   %s" % (i, ms, " | ".join(inside[:2] + inside[-1:])))
print("idle sample:", samples[0][1])
