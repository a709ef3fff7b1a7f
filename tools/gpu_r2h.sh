#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2h_pytest.txt 2>&1; echo "rc=$?" >> gpurun_out/r2h_pytest.txt
tail -n 5 gpurun_out/r2h_pytest.txt
timeout 600 compute-sanitizer --tool memcheck python tools/gpu_sanitize.py ring pcg > gpurun_out/r2h_sanitizer.txt 2>&1; tail -n 4 gpurun_out/r2h_sanitizer.txt
timeout 600 python tools/gpu_pull_window.py cfg2 0 512 > gpurun_out/r2h_pull_window_cfg2.txt 2>&1; cat gpurun_out/r2h_pull_window_cfg2.txt
timeout 600 python tools/gpu_pull_window.py cfg5 512 > gpurun_out/r2h_pull_window_cfg5.txt 2>&1; cat gpurun_out/r2h_pull_window_cfg5.txt
LFR_TOOL_FLAGS=0x8 timeout 600 python tools/gpu_pull_window.py cfg5 512 >> gpurun_out/r2h_pull_window_cfg5.txt 2>&1; tail -n 1 gpurun_out/r2h_pull_window_cfg5.txt
timeout 900 python tools/gpu_cfg5.py cfg5 > gpurun_out/r2h_cfg5.json 2> gpurun_out/r2h_cfg5.err; grep -E "host_stage_s|plan_create_s|solve_ms|deterministic" gpurun_out/r2h_cfg5.json; tail -n 3 gpurun_out/r2h_cfg5.err
