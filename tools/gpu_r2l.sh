#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/run_cfg4.py <<'PY'
import sys
sys.path.insert(0, '.')
from lfr_b200 import build_problem, synth
from lfr_b200.capi import Plan, load_b200
import torch
p = build_problem(synth.generate("cfg4"))
lib = load_b200()
plan = Plan(lib, p)
s = torch.cuda.current_stream().cuda_stream
plan.solve(s); torch.cuda.synchronize()
PY
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:solve_tile_kernel -c 4 -o gpurun_out/r2l_tile python /tmp/run_cfg4.py > gpurun_out/r2l_ncu.log 2>&1
tail -n 3 gpurun_out/r2l_ncu.log; ls -la gpurun_out/r2l_tile.ncu-rep
