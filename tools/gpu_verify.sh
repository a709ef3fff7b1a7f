#!/bin/bash
# last check of the round on one B200: the GPU test suite
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/r02_pytest_gpu.txt 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_gpu.txt; tail -n 3 gpurun_out/r02_pytest_gpu.txt
