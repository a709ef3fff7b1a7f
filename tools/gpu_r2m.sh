#!/bin/bash
mkdir -p gpurun_out
nproc
for t in 1 2 3 4 8; do echo "threads $t"; LFR_SCHEDULE_THREADS=$t python tools/host_schedule_time.py cfg2 cfg3 cfg4 2>&1 | tr '\n' ' '; echo; done | tee gpurun_out/r2m_schedule_threads.txt
echo default; python tools/host_schedule_time.py cfg2 cfg3 cfg4 cfg5 | tee -a gpurun_out/r2m_schedule_threads.txt
LFR_SCHEDULE_THREADS=1 python tools/host_schedule_time.py cfg5 | tee -a gpurun_out/r2m_schedule_threads.txt
