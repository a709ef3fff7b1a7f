#!/bin/bash
# round 2, first GPU pass: correctness of the staging / zero-copy paths, then numbers
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_gpu.txt 2>&1
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2a_smoke.txt 2>&1; echo "smoke rc=$?" >> gpurun_out/r2a_smoke.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "staging_and_zero_copy" > gpurun_out/r2a_zc.txt 2>&1; echo "rc=$?" >> gpurun_out/r2a_zc.txt
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2a_pytest.txt 2>&1; echo "rc=$?" >> gpurun_out/r2a_pytest.txt
for flags in 0 8 4 12; do
  LFR_BENCH_DEBUG=1 timeout 300 python bench.py --steps 20 --warmup 3 --debug-flags $flags > gpurun_out/r2a_bench_cfg2_f$flags.json 2> gpurun_out/r2a_bench_cfg2_f$flags.err
done
timeout 300 python bench.py --steps 10 --warmup 3 --workload cfg4 > gpurun_out/r2a_bench_cfg4.json 2> gpurun_out/r2a_bench_cfg4.err
timeout 300 python bench.py --steps 10 --warmup 3 --workload cfg3 > gpurun_out/r2a_bench_cfg3.json 2> gpurun_out/r2a_bench_cfg3.err
timeout 300 python tools/gpu_cycles.py cfg2 > gpurun_out/r2a_cycles_cfg2.txt 2>&1
tail -3 gpurun_out/r2a_smoke.txt gpurun_out/r2a_zc.txt gpurun_out/r2a_pytest.txt
for f in gpurun_out/r2a_bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], 'ms/step %.4f'%d['ms_per_step'], 'e2e ms %.4f'%d['e2e']['ms_per_step'], d['e2e']['stages_ms'], 'frac %.4f'%d['roofline']['frac'])
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
done
