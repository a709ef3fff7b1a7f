#!/bin/bash
# round 2, pass g: device-side CTA preparation + zero-copy pull pacing
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2g_pytest.txt 2>&1; echo "rc=$?" >> gpurun_out/r2g_pytest.txt
tail -n 5 gpurun_out/r2g_pytest.txt
timeout 600 python tools/gpu_pull_window.py cfg2 > gpurun_out/r2g_pull_window_cfg2.txt 2>&1; cat gpurun_out/r2g_pull_window_cfg2.txt
timeout 600 python tools/gpu_pull_window.py cfg4 0 512 2048 > gpurun_out/r2g_pull_window_cfg4.txt 2>&1; cat gpurun_out/r2g_pull_window_cfg4.txt
timeout 600 compute-sanitizer --tool memcheck python tools/gpu_sanitize.py ring pcg > gpurun_out/r2g_sanitizer.txt 2>&1; tail -n 4 gpurun_out/r2g_sanitizer.txt
timeout 600 compute-sanitizer --tool racecheck python tools/gpu_sanitize.py ring > gpurun_out/r2g_racecheck.txt 2>&1; tail -n 4 gpurun_out/r2g_racecheck.txt
timeout 900 python tools/gpu_cfg5.py cfg5 > gpurun_out/r2g_cfg5.json 2> gpurun_out/r2g_cfg5.err; cat gpurun_out/r2g_cfg5.json | head -c 1500; tail -n 3 gpurun_out/r2g_cfg5.err
