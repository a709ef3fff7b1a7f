#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -k "cta or CTA or large or ring or madrid or dense or pcg or zero_copy or malformed or multi_device" > gpurun_out/r2n_pytest_cta.txt 2>&1; tail -n 3 gpurun_out/r2n_pytest_cta.txt
timeout 600 python tools/gpu_pull_window.py cfg5 512 > gpurun_out/r2n_e2e_cfg5.txt 2>&1; cat gpurun_out/r2n_e2e_cfg5.txt
