#!/bin/bash
# One GPU-box visit (round 2): parity tests, smoke, bench lines, the ncu launch list of one steady-state env step and a
# full-set capture of the kernels of one substep.  Usage: gpurun --timeout 2400 -- 'bash tools/gpu_round2.sh TAG [full]'
TAG=${1:-r2}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke_$TAG.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_$TAG.json
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --envs-per-gpu 8192 2>&1 | tail -1 | tee gpurun_out/bench8k_$TAG.json
# launch list: one steady-state env step (the timed region starts the capture)
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:"prb_" -c 200 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --profiler-range > gpurun_out/b_launch_$TAG.log 2>&1
python tools/launch_summary.py gpurun_out/launches_$TAG.csv | tee gpurun_out/launch_summary_$TAG.txt
if [ "$2" = full ]; then
  # full-set capture of one substep's kernels (skip the IK launch and the first substeps)
  ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"prb_setup|prb_pgs" -s 16 -c 8 -f -o gpurun_out/prof_$TAG \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline --profiler-range > gpurun_out/b_ncu_$TAG.log 2>&1
  tail -c 300 gpurun_out/b_ncu_$TAG.log
fi
