#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python tools/gpu_cta_parity.py > gpurun_out/r2e_cta_parity.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "cta or pcg or fuzz" > gpurun_out/r2e_pytest_cta.txt 2>&1; echo "rc=$?" >> gpurun_out/r2e_pytest_cta.txt
timeout 900 python tools/gpu_cfg5.py cfg5 > gpurun_out/r2e_cfg5.txt 2>&1
cat gpurun_out/r2e_cta_parity.txt | cut -c1-600
tail -n 4 gpurun_out/r2e_pytest_cta.txt
grep -E "solve_ms|host_stage_s|gen_s|tracks_per_s|deterministic|cost_never" gpurun_out/r2e_cfg5.txt
