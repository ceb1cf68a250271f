TAG=$1
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_$TAG.json
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --jump-frac 0.05 2>&1 | tail -1 > gpurun_out/benchj_$TAG.json
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --envs-per-gpu 8192 2>&1 | tail -1 > gpurun_out/bench8k_$TAG.json
tail -c 300 gpurun_out/bench_$TAG.json
