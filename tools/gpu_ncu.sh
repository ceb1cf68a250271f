#!/bin/bash
# ncu --set full capture of selected kernels of a short bench run.  Usage: gpu_ncu.sh TAG REGEX SKIP COUNT [ENVS]
TAG=$1; RX=$2; SKIP=$3; CNT=$4; ENVS=${5:-8192}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"$RX" -s $SKIP -c $CNT -f -o gpurun_out/prof_$TAG \
    python bench.py --steps 3 --warmup 3 --envs-per-gpu $ENVS --no-cpu-baseline > gpurun_out/b_ncu_$TAG.log 2>&1
tail -c 300 gpurun_out/b_ncu_$TAG.log
