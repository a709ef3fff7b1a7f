#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -k "cta or CTA or large or ring or madrid or dense or pcg" > gpurun_out/r2i_pytest_cta.txt 2>&1; tail -n 3 gpurun_out/r2i_pytest_cta.txt
timeout 900 python tools/gpu_cfg5_minb.py cfg5 2,2,3,4//256,256,256,256 1,2,3,4//512,512,256,256 1,2,2,4//512,512,512,256 1,2,2,2//512,512,512,512 1,1,2,4//512,512,512,256 > gpurun_out/r2i_cfg5_minb.txt 2>&1; cat gpurun_out/r2i_cfg5_minb.txt | tail -n 12
