#!/bin/bash
# last check of the round on one B200: GPU tests, smoke, the bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r02_pytest_gpu.txt 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_gpu.txt; tail -n 3 gpurun_out/r02_pytest_gpu.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02_smoke.txt 2>&1; tail -n 1 gpurun_out/r02_smoke.txt
timeout 600 python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; tail -c 900 gpurun_out/r02_bench_default.json
