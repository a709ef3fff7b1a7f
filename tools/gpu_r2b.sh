#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2b_pytest.txt 2>&1; echo "rc=$?" >> gpurun_out/r2b_pytest.txt
for flags in 0 8; do
  LFR_BENCH_DEBUG=1 timeout 300 python bench.py --steps 20 --warmup 3 --debug-flags $flags > gpurun_out/r2b_bench_cfg2_f$flags.json 2> gpurun_out/r2b_bench_cfg2_f$flags.err
done
timeout 300 python bench.py --steps 10 --warmup 3 --workload cfg3 > gpurun_out/r2b_bench_cfg3.json 2> gpurun_out/r2b_bench_cfg3.err
timeout 300 python tools/gpu_cycles.py cfg2 > gpurun_out/r2b_cycles_cfg2.txt 2>&1
tail -n 3 gpurun_out/r2b_pytest.txt
head -12 gpurun_out/r2b_cycles_cfg2.txt
for f in gpurun_out/r2b_bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], 'ms/step %.4f'%d['ms_per_step'], 'e2e ms %.4f'%d['e2e']['ms_per_step'], d['e2e']['stages_ms'], 'frac %.4f'%d['roofline']['frac'])
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
done
timeout 300 python tools/gpu_zc_timeline.py cfg2 0 > gpurun_out/r2b_timeline_zc.txt 2>&1
timeout 300 python tools/gpu_zc_timeline.py cfg2 8 > gpurun_out/r2b_timeline_hbm.txt 2>&1
cat gpurun_out/r2b_timeline_zc.txt gpurun_out/r2b_timeline_hbm.txt
