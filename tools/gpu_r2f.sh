#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2f_pytest.txt 2>&1; echo "rc=$?" >> gpurun_out/r2f_pytest.txt
LFR_BENCH_DEBUG=1 timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2f_bench_cfg2.json 2> gpurun_out/r2f_bench_cfg2.err
timeout 300 python tools/gpu_cycles.py cfg2 > gpurun_out/r2f_cycles_cfg2.txt 2>&1
python local-feature-refinement_b200/csrc/build.py --poly-prof > /dev/null 2>&1 && timeout 300 python tools/gpu_polyprof.py cfg2 > gpurun_out/r2f_polyprof.txt 2>&1
tail -n 3 gpurun_out/r2f_pytest.txt; head -14 gpurun_out/r2f_cycles_cfg2.txt; cat gpurun_out/r2f_polyprof.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2f_bench_cfg2.json').read().strip().splitlines()[-1])
print('ms/step %.4f'%d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['stages_ms'])
PY
