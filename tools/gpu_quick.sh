#!/bin/bash
# Quick GPU check while iterating: parity tests + a short bench line (no CPU baseline leg).
TAG=${1:-q}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_$TAG.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_$TAG.json
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --envs-per-gpu 8192 2>&1 | tail -1 | tee gpurun_out/bench8k_$TAG.json
