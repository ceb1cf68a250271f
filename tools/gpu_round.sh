#!/bin/bash
# One GPU-box visit: parity tests, smoke, the bench line, the reference arm, the ncu launch list of steady-state
# env steps and a full-set capture of every kernel of the step pipeline at the bench size.
# Usage: gpurun --timeout 2400 -- 'bash tools/gpu_round.sh TAG'
TAG=${1:-r1}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke_$TAG.log
python bench.py --steps 20 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_$TAG.json
python bench.py --impl reference --steps 3 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_ref_$TAG.json
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --jump-frac 0.05 2>&1 | tail -1 > gpurun_out/bench_jump_$TAG.json
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --preroll 0 2>&1 | tail -1 > gpurun_out/bench_nopreroll_$TAG.json
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --envs-per-gpu 8192 2>&1 | tail -1 > gpurun_out/bench_8192_$TAG.json
for E in UR5Reach-v0 pandaPick-v0; do
  N=4096; [ $E = pandaPick-v0 ] && N=16384
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --env $E --envs-per-gpu $N 2>&1 | tail -1 > gpurun_out/bench_${E}_$TAG.json
done
# launch list: two steady-state env steps (after the 96-step pre-roll)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"prb_setup|prb_pgs|prb_ik" -s 6040 -c 130 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b_launch_$TAG.log 2>&1
# full-set capture of one substep's kernels (setup, arm-island x 2, joint, free) in the steady state
ncu --set full --clock-control none --import-source on -k regex:"prb_setup_kernel|prb_pgs" -s 6105 -c 6 -f -o gpurun_out/prof_$TAG \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu_$TAG.log 2>&1
tail -c 300 gpurun_out/b_ncu_$TAG.log
