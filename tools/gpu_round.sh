#!/bin/bash
# One GPU-box visit: parity tests, the bench line, the reference arm, the ncu launch list and a
# full capture of the step kernels.  Usage: gpurun --timeout 1500 -- 'bash tools/gpu_round.sh TAG'
TAG=${1:-r1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > gpurun_out/clocks_$TAG.csv &
SMI=$!
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke_$TAG.log
python bench.py --steps 20 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_$TAG.json
python bench.py --impl reference --steps 3 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_ref_$TAG.json
kill $SMI
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 3 --warmup 3 --envs-per-gpu 8192 --no-cpu-baseline > gpurun_out/b_launch_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"prb_setup_kernel|prb_pgs_kernel" -s 85 -c 4 -f -o gpurun_out/prof_$TAG \
    python bench.py --steps 3 --warmup 3 --envs-per-gpu 8192 --no-cpu-baseline > gpurun_out/b_ncu_$TAG.log 2>&1
tail -c 600 gpurun_out/b_ncu_$TAG.log
# the same two kernels at the bench size (65536 envs: the record stream no longer fits L2)
ncu --set full --clock-control none --import-source on -k regex:"prb_setup_kernel|prb_pgs_kernel" -s 85 -c 2 -f -o gpurun_out/prof64k_$TAG \
    python bench.py --steps 3 --warmup 3 --envs-per-gpu 65536 --no-cpu-baseline > gpurun_out/b_ncu64k_$TAG.log 2>&1
tail -c 300 gpurun_out/b_ncu64k_$TAG.log
