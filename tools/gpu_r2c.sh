#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2c_pytest.txt 2>&1; echo "rc=$?" >> gpurun_out/r2c_pytest.txt
LFR_BENCH_DEBUG=1 timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2c_bench_cfg2.json 2> gpurun_out/r2c_bench_cfg2.err
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r2c_bench_cfg2_reference.json 2> gpurun_out/r2c_bench_cfg2_reference.err
timeout 300 python tools/gpu_cycles.py cfg2 > gpurun_out/r2c_cycles_cfg2.txt 2>&1
timeout 600 python tools/gpu_batched.py > gpurun_out/r2c_batched.txt 2>&1
tail -n 3 gpurun_out/r2c_pytest.txt
head -14 gpurun_out/r2c_cycles_cfg2.txt
tail -n 8 gpurun_out/r2c_batched.txt
python - <<'PY'
import json
for f in ('gpurun_out/r2c_bench_cfg2.json','gpurun_out/r2c_bench_cfg2_reference.json'):
    try:
        d=json.load(open(f))
        print(f, 'ms/step %.4f'%d['ms_per_step'], 'e2e', d['e2e'].get('ms_per_step'), d['e2e'].get('stages_ms'), 'cpu', d['cpu_baseline'].get('cores'), d['cpu_baseline'].get('ms_per_step'), d['cpu_baseline'].get('eight_thread_ms_per_step'), d['cpu_baseline'].get('eight_thread_ms_per_scene'), d.get('total_scope'))
    except Exception as e:
        print(f,'ERR',e)
PY
