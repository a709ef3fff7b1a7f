#!/bin/bash
# multi-GPU pass: tests, sharded record through the driver's launch line
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/multi_gpus.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "multi_device or launcher_with_gpus" > gpurun_out/multi_pytest_multi.txt 2>&1; echo "rc=$?" >> gpurun_out/multi_pytest_multi.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/multi_bench_${N}gpu.json 2> gpurun_out/multi_bench_${N}gpu.err
tail -n 5 gpurun_out/multi_pytest_multi.txt
tail -n 5 gpurun_out/multi_bench_${N}gpu.err
python - $N <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/multi_bench_%sgpu.json'%n).read().strip().splitlines()[-1])
    print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])
    print(json.dumps(d.get('sharded'), indent=1))
except Exception as e:
    print('ERR', e)
PY
