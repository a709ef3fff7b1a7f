#!/bin/bash
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:solve_cta_kernel -c 3 -o gpurun_out/r2j_cta python tools/gpu_cfg5.py cfg5 > gpurun_out/r2j_ncu.log 2>&1
tail -n 5 gpurun_out/r2j_ncu.log; ls -la gpurun_out/r2j_cta.ncu-rep
