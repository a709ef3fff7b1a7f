#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -k "native" > gpurun_out/r02_pytest_native.txt 2>&1; tail -n 2 gpurun_out/r02_pytest_native.txt
python - <<'PY' | tee gpurun_out/r02_native_exe_wall.txt
import sys, subprocess, time, os
sys.path.insert(0, '.')
from lfr_b200 import synth, wire
ms = synth.generate('cfg2')
open('/tmp/cfg2.pb', 'wb').write(wire.encode_matching_file(ms))
for exe in (['multi-view-refinement/build/solve_native'], [sys.executable, 'multi-view-refinement/build/solve']):
    best = None
    for rep in range(3):
        t = time.perf_counter(); r = subprocess.run(exe + ['--matches_file', '/tmp/cfg2.pb', '--output_file', '/tmp/o_%d.pb' % len(exe)], capture_output=True, text=True); dt = time.perf_counter() - t
        best = dt if best is None else min(best, dt)
    print(exe[-1], 'wall %.3f s (best of 3)' % best, '|', ' | '.join(l for l in r.stdout.splitlines() if 'time' in l), r.stderr[-200:])
print('same output:', open('/tmp/o_1.pb','rb').read() == open('/tmp/o_2.pb','rb').read())
env = dict(os.environ, CUDA_VISIBLE_DEVICES='0,1,2,3,4,5,6,7')
t = time.perf_counter(); subprocess.run(['multi-view-refinement/build/solve_native', '--matches_file', '/tmp/cfg2.pb', '--output_file', '/tmp/o3.pb'], capture_output=True, env=env); print('native, all GPUs visible: wall %.3f s' % (time.perf_counter() - t))
PY
