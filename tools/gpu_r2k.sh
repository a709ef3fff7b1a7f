#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/gpu_tier_timeline.py cfg4 > gpurun_out/r2k_timeline_cfg4.txt 2>&1; cat gpurun_out/r2k_timeline_cfg4.txt | head -n 12
for w in cfg2 cfg4; do timeout 600 python bench.py --workload $w --steps 20 --warmup 3 > gpurun_out/r2k_bench_$w.json 2> gpurun_out/r2k_bench_$w.err; python - $w <<'PY'
import json,sys
w=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/r2k_bench_%s.json'%w).read().strip().splitlines()[-1])
    print(w,'ms/step %.4f'%d['ms_per_step'],'e2e %.4f'%d['e2e']['ms_per_step'],d['e2e'].get('stages_ms'))
except Exception as e: print(w,'ERR',e)
PY
done
