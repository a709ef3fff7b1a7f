#!/bin/bash
# End-of-round evidence on one B200: tests, smoke, bench lines, ncu launch list and full capture.
# Usage (from the repo root on the GPU box): bash tools/gpu_round_report.sh r01
set -u
R=${1:-r01}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q 2>&1 | tail -3 > $O/${R}_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/${R}_smoke.txt 2>&1
for w in cfg2 cfg3 cfg4; do
  python bench.py --steps 20 --warmup 5 --workload $w > $O/${R}_bench_$w.json 2> $O/${R}_bench_$w.err
done
python bench.py --impl reference --steps 5 --warmup 3 > $O/${R}_bench_cfg2_reference.json 2> $O/${R}_bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${R}_launches.csv \
    python bench.py --steps 2 --warmup 3 > $O/${R}_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:solve_ -o $O/${R}_full_cfg2 -f \
    python tools/gpu_prof.py cfg2 1 > $O/${R}_full_cfg2.log 2>&1
python tools/gpu_cycles.py cfg2 > $O/${R}_cycles_cfg2.txt 2>&1
python tools/gpu_batched.py > $O/${R}_batched.txt 2>&1
tail -2 $O/${R}_pytest_gpu.txt; cat $O/${R}_smoke.txt | tail -1
for w in cfg2 cfg3 cfg4; do python - <<PY
import json
d=json.loads(open("$O/${R}_bench_$w.json").read().strip().splitlines()[-1])
print("$w", round(d["ms_per_step"],4), round(d["value"]), "e2e", round(d["e2e"]["ms_per_step"],4), d["roofline"]["frac"], d["clocks"])
PY
done
tail -c 600 $O/${R}_bench_cfg2_reference.json
