"""lfr_solve_multi on device subsets, zero-copy vs through HBM (single process, no torchrun)."""
import ctypes as C
import os
import sys
import time

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from lfr_b200 import build_problem, synth  # noqa: E402
from lfr_b200.capi import LfrMultiInfo, load_b200  # noqa: E402

lib = load_b200()
name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
p = build_problem(synth.generate(name))
s2, keep, pos_pinned, h2d = bench.pinned_problem(lib, p)
stt, bufs = lib.make_stats(p.n_components)
nd = torch.cuda.device_count()
sets = [[d] for d in range(nd)] + ([list(range(nd))] if nd > 1 else [])
if nd >= 4:
    sets += [[0, 1], [2, 3]]
for flags in (0, 8):
    opts = lib.default_options(debug_flags=flags)
    for devices in sets:
        dev = np.ascontiguousarray(devices, dtype=np.int32)
        info = LfrMultiInfo()
        wall = []
        for i in range(6):
            pos_pinned.zero_()
            t1 = time.perf_counter()
            rc = lib.lib.lfr_solve_multi(C.byref(s2), C.byref(opts), dev.ctypes.data, len(devices), pos_pinned.data_ptr(),
                                         C.byref(stt), C.byref(info))
            dt = time.perf_counter() - t1
            lib.check(rc, "multi")
            if i >= 2:
                wall.append(dt)
        print("flags", flags, "devices", devices, "wall ms %.3f" % (1e3 * float(np.median(wall))),
              "kernel ms", [round(x, 3) for x in list(info.kernel_ms)[:len(devices)]],
              "total ms", [round(x, 3) for x in list(info.total_ms)[:len(devices)]], flush=True)
