#!/bin/bash
# Quick GPU check while iterating: parity tests + a short bench line (no CPU baseline leg) + per-kernel
# launch durations of one env step at the bench size.
TAG=${1:-q}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_$TAG.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_$TAG.json
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --envs-per-gpu 8192 2>&1 | tail -1 | tee gpurun_out/bench8k_$TAG.json
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"prb_setup|prb_pgs|prb_ik" -s 200 -c 70 --csv --log-file gpurun_out/launches64k_$TAG.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b_l64k_$TAG.log 2>&1
python - <<PY
import csv, collections
d = collections.defaultdict(list)
for r in csv.DictReader(l for l in open('gpurun_out/launches64k_$TAG.csv') if l.startswith('"')):
    if r.get('Metric Name') == 'gpu__time_duration.sum':
        d[r['Kernel Name'].split('(')[0]].append(float(r['Metric Value'].replace(',', '')) / 1e6)
for k, v in d.items():
    print('%-40s n=%3d mean %.3f ms min %.3f max %.3f' % (k[:40], len(v), sum(v) / len(v), min(v), max(v)))
PY
